"""The knowledge-distillation generator step (reference train.py:280-308, `G_Loss_BackProp`) as a
reusable driver: student forward (RGB list) -> discriminator -> non-saturating GAN loss; teacher
forward; content mask (train.py:154-158); masked L1 + LPIPS KD losses (train.py:145-184); backward; one flat-bucket
gradient all-reduce over the data-parallel ranks; fused Adam (train.py:528-532 hyper-parameters).

The LPIPS-VGG16 distance runs on the package's own kernels (b200gan/lpips.PerceptualLossVGG; any callable with
`lpips.PerceptualLoss`'s contract is accepted); the face parser (BiSeNet) is a caller-supplied module
with BiSeNet's call contract -- what goes in and what comes out of it is the device-side glue of b200gan/maskglue.py.
Without a parser a constant content mask can be given.  Public API used by bench.py's e2e leg.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import dist as D
from . import maskglue
from ._lib import lib, check


class KDStep:
    def __init__(self, student, teacher, discriminator, lr: float = 0.002 * 0.8, betas=(0.0, 0.99 ** 0.8),
                 eps: float = 1e-8, kd_l1_lambda: float = 3.0, mask: Optional[torch.Tensor] = None,
                 kd_mode: str = 'Output_Only', g_ema=None, ema_decay: float = 0.5 ** (32 / (10 * 1000)),
                 percept_loss=None, kd_lpips_lambda: float = 3.0, parsing_net=None, lpips_image_size: int = 256):
        self.student, self.teacher, self.disc = student, teacher, discriminator
        # train.py:509-520: percept_loss(pred, target) -> [N,1,1,1] (LPIPS-VGG), parsing_net(x)[0] -> class scores
        self.percept_loss, self.kd_lpips_lambda = percept_loss, float(kd_lpips_lambda)
        self.parsing_net, self.lpips_image_size = parsing_net, int(lpips_image_size)
        for p in teacher.parameters():
            p.requires_grad_(False)
        for p in discriminator.parameters():
            p.requires_grad_(False)          # train.py:286-287 (requires_grad(D, False))
        # the frozen discriminator runs on the package's own engines (b200gan/dconv.py): NHWC-p activations, weights
        # re-laid out once into operand slabs (cached: the parameters are frozen), the image read through its strides
        self.bucket = D.FlatBucket(student.parameters())
        # exponential moving average of the student (train.py:124-129 / :398; default decay train.py:352): kept in a
        # bucket with the student's layout and updated by the same kernel pass as Adam
        self.ema_bucket = None
        if g_ema is not None:
            names = [n for n, p in student.named_parameters() if p.requires_grad]
            ema_params = dict(g_ema.named_parameters())
            self.ema_bucket = D.FlatBucket([ema_params[n] for n in names], with_grads=False, all_params=True)
            assert self.ema_bucket.offsets == self.bucket.offsets, 'g_ema must mirror the student parameter by parameter'
        self.g_ema, self.ema_decay = g_ema, float(ema_decay)
        self.exp_avg = torch.zeros_like(self.bucket.flat_param)
        self.exp_avg_sq = torch.zeros_like(self.bucket.flat_param)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.kd_l1_lambda = kd_l1_lambda
        self.mask = mask
        if kd_mode not in ('Output_Only', 'Intermediate'):
            raise ValueError(f'kd_mode {kd_mode!r}: expected Output_Only or Intermediate (train.py:163-169)')
        self.kd_mode = kd_mode
        self.device = self.bucket.flat_param.device
        self.t_dev = torch.zeros(1, device=self.device)      # Adam step count, device-resident (graph-safe)
        self.graph = None
        self.teacher_stream = torch.cuda.Stream(device=self.device) \
            if (self.device.type == 'cuda' and os.environ.get('CAGC_KD_OVERLAP', '1') != '0') else None

    def losses(self, z: List[torch.Tensor], inject_index: int, s_noise=None, t_noise=None):
        """GAN + KD losses; per-layer noise is drawn fresh (train.py:291,151) unless given explicitly."""
        # The frozen teacher does not depend on the student: its forward runs on a side stream (a parallel
        # branch of the captured graph) that only waits for the latents, and fills the SMs that the student's
        # small low-resolution launches leave idle.  Program order stays student -> teacher, so the per-layer
        # noise draws consume the RNG stream in the reference's order (train.py:291, then :151).
        main = torch.cuda.current_stream(self.device)
        ready = main.record_event() if self.teacher_stream is not None else None
        fake = self.student(z, return_rgb_list=True, inject_index=inject_index, noise=s_noise)
        mask = self.mask
        if self.teacher_stream is not None:
            self.teacher_stream.wait_event(ready)
            with torch.cuda.stream(self.teacher_stream), torch.no_grad():
                real = self.teacher(z, return_rgb_list=True, inject_index=inject_index, noise=t_noise)
                for r in real:
                    r.record_stream(main)
                if self.parsing_net is not None:     # the face parser only needs the teacher image: same branch
                    mask = maskglue.content_mask(real[-1], self.parsing_net)
                    mask.record_stream(main)
        else:
            with torch.no_grad():
                real = self.teacher(z, return_rgb_list=True, inject_index=inject_index, noise=t_noise)
                if self.parsing_net is not None:
                    mask = maskglue.content_mask(real[-1], self.parsing_net)
        g_loss = F.softplus(-self.disc(fake[-1])).mean()
        if self.teacher_stream is not None:
            main.wait_stream(self.teacher_stream)
        # content-aware adjustment (train.py:154-158): the mask comes from parsing the TEACHER image and multiplies both
        s_img, t_img = fake[-1], real[-1]
        if mask is not None:
            s_img, t_img = s_img * mask, t_img * mask
        if self.kd_mode == 'Output_Only':                      # train.py:163-164
            kd = self.kd_l1_lambda * torch.mean(torch.abs(t_img - s_img))
        else:
            # train.py:165-169: every resolution of the two rgb lists.  The list comprehension has its own scope, so the L1
            # terms pair the UNMASKED list entries; the `for fake_img_teacher in fake_img_teacher_list` loop before it,
            # however, leaves `fake_img_teacher` bound to the last (unmasked) teacher image, which is what LPIPS then sees
            kd = self.kd_l1_lambda * sum(torch.mean(torch.abs(t - f)) for t, f in zip(real, fake))
            t_img = real[-1]
        if self.percept_loss is not None:                      # train.py:172-182
            p_s, p_t = s_img, t_img
            if s_img.shape[-1] > self.lpips_image_size:         # "pooled the image for LPIPS for memory saving"
                size = (self.lpips_image_size, self.lpips_image_size)
                p_s = F.interpolate(s_img, size=size, mode='bilinear', align_corners=False)
                p_t = F.interpolate(t_img, size=size, mode='bilinear', align_corners=False)
            kd = kd + self.kd_lpips_lambda * torch.mean(self.percept_loss(p_s, p_t))
        return g_loss, kd

    def step(self, z: List[torch.Tensor], inject_index: int, s_noise=None, t_noise=None) -> torch.Tensor:
        """One optimisation step on device-resident latents; returns the (detached) total loss."""
        self.bucket.detach_grads()
        g_loss, kd = self.losses(z, inject_index, s_noise, t_noise)
        total = g_loss + kd
        total.backward()
        self.bucket.pack_grads()
        self.bucket.allreduce_mean_()
        self.t_dev += 1
        b1, b2 = self.betas
        with torch.cuda.device(self.device):
            check(lib.cagc_adam_ema_step(torch.cuda.current_stream(self.device).cuda_stream,
                                         self.bucket.flat_param.data_ptr(), self.bucket.flat_grad.data_ptr(),
                                         self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.bucket.numel,
                                         self.lr, b1, b2, self.eps, 1.0 / D.get_world_size(), 1.0, 1.0,
                                         self.t_dev.data_ptr(),
                                         None if self.ema_bucket is None else self.ema_bucket.flat_param.data_ptr(),
                                         self.ema_decay), 'adam_step')
        self._bump_versions()
        return total.detach()

    def _bump_versions(self):
        """The fused Adam kernel writes the parameters behind autograd's back (raw pointer into the flat bucket):
        bump every parameter's version counter so that anything keyed on it -- the frozen-weight operand cache of
        b200gan.modconv when the student is frozen for a discriminator step (train.py:247), autograd's saved-tensor
        checks -- sees the update."""
        torch._C._increment_version(self.bucket.params)
        if self.ema_bucket is not None:
            torch._C._increment_version(self.ema_bucket.params)

    # ------------------------------------------------------------------ CUDA-graph form
    def capture(self, batch: int, inject_index: int, style_dim: int = 512):
        """Capture the whole step (forward, backward, all-reduce, Adam) into one CUDA graph: a few
        thousand launches per step become one graph launch.  Shapes, inject_index and the set of
        tensors are frozen; latents are fed through static input buffers; per-layer noise is drawn
        inside the graph by torch's graph-safe Philox generator."""
        dev = self.device
        self.z_static = [torch.zeros(batch, style_dim, device=dev), torch.zeros(batch, style_dim, device=dev)]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                self.z_static[0].normal_()
                self.z_static[1].normal_()
                self.step(self.z_static, inject_index)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        from ._lib import launch_count
        n0 = launch_count()
        with torch.cuda.graph(self.graph):
            self.loss_static = self.step(self.z_static, inject_index)
        self.launches_per_replay = launch_count() - n0
        return self

    def step_graphed(self, z: List[torch.Tensor]) -> torch.Tensor:
        """Replay the captured step on new latents (device or pinned-host tensors)."""
        for dst, src in zip(self.z_static, z):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        self._bump_versions()          # a replay runs the Adam kernel too; the capture-time bump is not replayed
        return self.loss_static

    def step_from_host(self, z_host: List[torch.Tensor], inject_index: int) -> float:
        """End-to-end form: latents arrive in pinned host memory, the loss is read back."""
        if self.graph is not None:
            return float(self.step_graphed(z_host).item())
        z = [t.to(self.device, non_blocking=True) for t in z_host]
        return float(self.step(z, inject_index).item())
