"""Content-mask glue of the KD loss on the device.

Reference: train.py:154-158 -> `Batch_Img_Parsing` (Util/content_aware_pruning.py:61-88) and `Get_Masked_Tensor`
(:90-117).  The reference normalises / resizes with seven ATen launches, takes an int64 argmax, builds the mask through
HOST tensors (`.type(torch.FloatTensor)`: a blocking device->host copy in the middle of the step, which also makes the
step impossible to capture in a CUDA graph) and multiplies image by image in a Python loop.  Here: one kernel in front of
the parser, one behind it, one broadcast multiply; nothing leaves the device.
"""
from __future__ import annotations

import torch

from ._lib import lib, check, stream_of, require_cuda

PARSING_SIZE = 512          # Util/content_aware_pruning.py:73,101


def parse_preprocess(img: torch.Tensor, parsing_size: int = PARSING_SIZE, channels_last: bool = False) -> torch.Tensor:
    """[N,3,S,S] generator output in [-1,1] -> the parser's input [N,3,P,P] (no gradient: :84-85 runs under no_grad and
    the teacher image carries none); channels_last: the same tensor in torch.channels_last memory format (what cuDNN's
    tensor-core convolutions read without a layout copy)."""
    require_cuda(img, 'parse_preprocess')
    n, c, s, s2 = img.shape
    if c != 3 or s != s2:
        raise RuntimeError(f'parse_preprocess: [N,3,S,S] image expected, got {tuple(img.shape)}')
    out = torch.empty((n, 3, parsing_size, parsing_size), device=img.device, dtype=torch.float32,
                      memory_format=torch.channels_last if channels_last else torch.contiguous_format)
    x = img.detach()
    sb, sc, sh, sw = x.stride()
    with torch.cuda.device(img.device):
        check(lib.cagc_parse_preprocess(stream_of(x), x.data_ptr(), sb, sc, sh, sw, out.data_ptr(), n, s, parsing_size,
                                        int(channels_last)), 'parse_preprocess')
    return out


def parsing_mask(logits: torch.Tensor, size: int) -> torch.Tensor:
    """Parser class scores [N,K,P,P] -> content mask [N,1,size,size] of 0/1 floats (the `resized_mask` of :103-109)."""
    require_cuda(logits, 'parsing_mask')
    n, k, p, p2 = logits.shape
    if p != p2:
        raise RuntimeError(f'parsing_mask: square class maps expected, got {tuple(logits.shape)}')
    lg = logits.detach().contiguous()
    mask = torch.empty((n, 1, size, size), device=logits.device, dtype=torch.float32)
    with torch.cuda.device(logits.device):
        check(lib.cagc_parsing_mask(stream_of(lg), lg.data_ptr(), mask.data_ptr(), n, k, p, size), 'parsing_mask')
    return mask


def parsing_mask_lowres(scores: torch.Tensor, size: int, parsing_size: int = PARSING_SIZE) -> torch.Tensor:
    """Low-resolution class scores [N,K,h,w] (before BiSeNet's final align_corners=True upsample to parsing_size) ->
    content mask [N,1,size,size]; any memory layout."""
    require_cuda(scores, 'parsing_mask_lowres')
    n, k, h, w = scores.shape
    sc = scores.detach()
    mask = torch.empty((n, 1, size, size), device=scores.device, dtype=torch.float32)
    sn, sk, sh, sw = sc.stride()
    with torch.cuda.device(scores.device):
        check(lib.cagc_parsing_mask_lowres(stream_of(sc), sc.data_ptr(), sn, sk, sh, sw, mask.data_ptr(), n, k, h, w,
                                           parsing_size, size), 'parsing_mask_lowres')
    return mask


def content_mask(teacher_img: torch.Tensor, parsing_net, parsing_size: int = PARSING_SIZE) -> torch.Tensor:
    """`Batch_Img_Parsing` + the mask half of `Get_Masked_Tensor`: [N,1,S,S]; `parsing_net(x)[0]` are the class scores
    (BiSeNet's call contract, Util/face_parsing/BiSeNet.py:230-254).  A parser that offers `scores_lowres(x)` (b200gan.
    parsing.FaceParser) is asked for the scores before its final upsample, which then happens inside the mask kernel."""
    with torch.no_grad():
        own = hasattr(parsing_net, 'scores_lowres')
        x = parse_preprocess(teacher_img, parsing_size, channels_last=own)
        if own:
            return parsing_mask_lowres(parsing_net.scores_lowres(x).float(), teacher_img.shape[-1], parsing_size)
        return parsing_mask(parsing_net(x)[0].float(), teacher_img.shape[-1])
