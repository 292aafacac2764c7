"""CPU restatement (torch, fp64-capable) of the rest of the KD loss: the LPIPS-VGG16 distance and the content-mask glue.

TEST INFRASTRUCTURE ONLY: imported by tests/, `__graft_entry__.smoke()` and bench.py's reference legs, never by the
product package.  Pinned against outputs of the reference's own `lpips` package and `Util/content_aware_pruning.py`
(tests/golden/lpips_tiny.npz, mask_glue.npz; tests/golden/make_golden.py `lpips` / `mask_glue`).
"""
import torch
import torch.nn.functional as F

# torchvision vgg16().features[0:30] cut into five slices (lpips/pretrained_networks.py:100-114)
VGG16_CFG = (64, 64, 'M', 128, 128, 'M', 256, 256, 256, 'M', 512, 512, 512, 'M', 512, 512, 512)
TAPS = (1, 3, 6, 9, 12)
SHIFT = (-.030, -.088, -.188)     # lpips/networks_basic.py:97-98
SCALE = (.458, .448, .450)


def vgg16_taps(x, conv_w, conv_b):
    """relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 of lpips/pretrained_networks.py:121-137."""
    taps, i = [], 0
    for c in VGG16_CFG:
        if c == 'M':
            x = F.max_pool2d(x, 2, 2)
            continue
        x = F.relu(F.conv2d(x, conv_w[i], conv_b[i], padding=1))
        if i in TAPS:
            taps.append(x)
        i += 1
    return taps


def normalize_tensor(x, eps=1e-10):
    """lpips/__init__.py:43-45."""
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def lpips_vgg(pred, target, conv_w, conv_b, lin_w):
    """`PerceptualLoss.forward(pred, target)` (lpips/__init__.py:27-41) = `PNetLin.forward(target, pred)`
    (lpips/networks_basic.py:62-92), net-lin, non-spatial, eval mode (dropout = identity): [N,1,1,1]."""
    shift = torch.tensor(SHIFT, dtype=pred.dtype).view(1, 3, 1, 1)
    scale = torch.tensor(SCALE, dtype=pred.dtype).view(1, 3, 1, 1)
    f0 = vgg16_taps((target - shift) / scale, conv_w, conv_b)
    f1 = vgg16_taps((pred - shift) / scale, conv_w, conv_b)
    val = 0
    for kk in range(5):
        d = (normalize_tensor(f0[kk]) - normalize_tensor(f1[kk])) ** 2
        val = val + F.conv2d(d, lin_w[kk].reshape(1, -1, 1, 1)).mean([2, 3], keepdim=True)
    return val


def batch_img_preprocess(img, parsing_size=512):
    """The tensor `Batch_Img_Parsing` feeds to the face parser (Util/content_aware_pruning.py:71-82)."""
    mean = torch.tensor([0.485, 0.456, 0.406], dtype=img.dtype).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], dtype=img.dtype).view(1, 3, 1, 1)
    t = ((img + 1) / 2).clamp(0, 1)
    t = F.interpolate(t, scale_factor=parsing_size / img.shape[-1], mode='bilinear', align_corners=False)
    return (t - mean) / std


def content_mask(parsing, size, parsing_size=512):
    """`Get_Masked_Tensor`'s mask (Util/content_aware_pruning.py:103-109): classes other than background (0) and
    class 16, resized bilinearly to the image size and thresholded at 0.5.  parsing: [N,P,P] integer labels."""
    m = ((parsing > 0) * (parsing != 16)).unsqueeze(0).double()
    r = F.interpolate(m, scale_factor=size / parsing_size, mode='bilinear', align_corners=False)
    return (r.squeeze(0) > 0.5).double()            # [N,size,size]


def get_masked_tensor(img, parsing, parsing_size=512):
    """Util/content_aware_pruning.py:90-117."""
    return img * content_mask(parsing, img.shape[-1], parsing_size).to(img.dtype).unsqueeze(1)
