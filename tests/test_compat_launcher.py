"""The reference's driver scripts, UNMODIFIED (byte-identical copies staged under oracle/_ref by
oracle/stage_ref.py), executed through the compat launcher against the drop-in `model` / `op` modules
(VERDICT r1 item 7).

  * prune.py: the pruned checkpoint it writes must contain exactly the channels the fp64 oracle selects for the
    same RNG streams (bit-exact prune mask);
  * get_fid.py / get_ppl.py: run end to end (sampling loops of Evaluation/fid.py:19-38, Evaluation/ppl.py) --
    Inception / VGG are random-init (no network), so only completion and finiteness are checked.
"""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'content-aware-gan-compression_b200')
REF = os.path.join(ROOT, 'oracle', '_ref')

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'prune.py')),
                                 reason='reference not staged (python oracle/stage_ref.py)')]

SHAPE = [48, 48, 48, 48, 48, 48, 40, 40, 32, 32]
SIZE = 64


def _launch(args, cwd, timeout=900):
    env = dict(os.environ)
    env['PYTHONPATH'] = PKG + os.pathsep + env.get('PYTHONPATH', '')
    env.pop('CAGC_CONV_ALGO', None)                   # the product default (tcgen05 TF32), as a user would run it
    r = subprocess.run([sys.executable, '-m', 'b200gan.run'] + args, cwd=cwd, env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert r.returncode == 0, f'launcher failed:\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}'
    return r.stdout, r.stderr


def _checkpoint(tmp_path):
    import model
    gen = synth.load_synth(model.Generator(SIZE, 512, 8, generator_net_shape=SHAPE), 61)
    sd = gen.state_dict()
    path = os.path.join(tmp_path, 'ckpt.pt')
    torch.save({'g': sd, 'g_ema': sd, 'd': {}}, path)
    return path, {k: v.double() for k, v in sd.items()}


def test_unmodified_prune_py_matches_oracle_masks(tmp_path):
    import model
    from PIL import Image
    from b200gan import saliency as S
    from oracle import stylegan2_oracle as O
    tmp = str(tmp_path)
    ckpt, sd64 = _checkpoint(tmp)
    os.makedirs(os.path.join(tmp, 'Model', 'pruned_model'))
    seed, n_sample, bs, prob, ratio = 5, 6, 3, 0.2, 0.5
    out, err = _launch(['--seed', str(seed), '--synthetic-mask', '--cwd', tmp, os.path.join(REF, 'prune.py'),
                        '--generated_img_size', str(SIZE), '--ckpt', ckpt, '--n_sample', str(n_sample),
                        '--batch_size', str(bs), '--noise_prob', str(prob), '--remove_ratio', str(ratio)], tmp)
    assert 'content-aware metric scoring takes' in out
    files = glob.glob(os.path.join(tmp, 'Model', 'pruned_model', '*.pth'))
    assert len(files) == 1
    pruned = torch.load(files[0], map_location='cpu')
    assert set(pruned) == {'g', 'd', 'g_ema'}

    # the oracle on the same streams: seeding, then the generator construction of Build_Generator_From_Dict (CPU
    # draws), then per batch: CPU latents, CUDA per-layer noise in execution order, numpy salt & pepper
    torch.manual_seed(seed)
    np.random.seed(seed)
    probe = model.Generator(SIZE, 512, 8, generator_net_shape=SHAPE)
    shapes = [(n.shape[2], n.shape[3]) for n in probe.make_noise()]
    torch.manual_seed(seed)                          # make_noise() above consumed CUDA/CPU draws: restart both streams
    model.Generator(SIZE, 512, 8, generator_net_shape=SHAPE)
    mask = np.array(Image.fromarray(S.default_mask(torch.zeros(1, 3, 512, 512)) > 0).resize((SIZE, SIZE)))
    per_batch = []
    for b in S.batch_sizes(n_sample, bs):
        z = torch.randn(b, 512)
        noise = [torch.empty(b, 1, h, w, device='cuda').normal_().double().cpu() for (h, w) in shapes]

        def noisy_fn(img):
            return S.noisy_images(img, lambda _i: mask, prob, np.random)
        per_batch.append(O.saliency_scores(sd64, SIZE, z.double(), noise, noisy_fn))
    total = S.total_scores(per_batch)
    masks = O.prune_mask_from_scores(total, ratio)
    sys.path.insert(0, REF)
    try:
        from Util.mask_util import Mask_the_Generator
        from Util.network_util import Get_Network_Shape
    finally:
        sys.path.remove(REF)
    expect = Mask_the_Generator({k: v.float() for k, v in sd64.items()}, masks)
    assert Get_Network_Shape(pruned['g_ema']) == [int(m.sum()) for m in masks]
    for k, v in expect.items():
        assert torch.equal(pruned['g_ema'][k].cpu(), v), f'pruned checkpoint differs at {k}'


def test_unmodified_get_fid_and_get_ppl_run(tmp_path):
    tmp = str(tmp_path)
    import model
    gen = synth.load_synth(model.Generator(256, 512, 8, generator_net_shape=[64] * 10 + [48, 48, 40, 40]), 62)
    ckpt = os.path.join(tmp, 'ckpt256.pt')
    torch.save({'g_ema': gen.state_dict()}, ckpt)
    out, _ = _launch(['--seed', '1', '--synthetic-fid-stats', os.path.join(REF, 'get_fid.py'), '--ckpt', ckpt,
                      '--n_sample', '32', '--batch_size', '16'], tmp)
    fid_line = [ln for ln in out.splitlines() if ln.startswith('FID Scores:')]
    assert fid_line and np.isfinite(float(fid_line[0].split(':')[1])), out[-500:]
    out, _ = _launch(['--seed', '1', os.path.join(REF, 'get_ppl.py'), '--ckpt', ckpt, '--n_sample', '16',
                      '--batch_size', '8'], tmp)
    ppl_line = [ln for ln in out.splitlines() if ln.startswith('PPL Scores:')]
    assert ppl_line and np.isfinite(float(ppl_line[0].split(':')[1])), out[-500:]


def test_unmodified_train_py_two_iterations(tmp_path):
    """train.py (D step, R1 with double backward at iteration 0, G/KD step with the content mask glue and LPIPS,
    path-length regulariser at iteration 0, EMA, sample grid) for two iterations on a synthetic image folder and
    synthetic checkpoints; VGG16 / the parser are random-init / synthetic (no network)."""
    import model
    from PIL import Image
    tmp = str(tmp_path)
    size, shape = 64, [48, 48, 48, 48, 48, 48, 40, 40, 32, 32]
    student = synth.load_synth(model.Generator(size, 512, 8, generator_net_shape=shape), 71)
    teacher = synth.load_synth(model.Generator(size, 512, 8, generator_net_shape=[64] * 10), 72)
    disc = synth.load_synth(model.Discriminator(size), 73)
    torch.save({'g': student.state_dict(), 'g_ema': student.state_dict(), 'd': disc.state_dict()}, os.path.join(tmp, 'student.pt'))
    torch.save({'g_ema': teacher.state_dict()}, os.path.join(tmp, 'teacher.pt'))
    data = os.path.join(tmp, 'data')
    os.makedirs(data)
    rs = np.random.RandomState(0)
    for i in range(8):
        Image.fromarray(rs.randint(0, 255, (size, size, 3), dtype=np.uint8)).save(os.path.join(data, f'{i:03d}.png'))
    out, err = _launch(['--seed', '3', '--synthetic-mask', '--cwd', tmp, os.path.join(REF, 'train.py'), '--path', data,
                        '--size', str(size), '--ckpt', os.path.join(tmp, 'student.pt'), '--teacher_ckpt',
                        os.path.join(tmp, 'teacher.pt'), '--iter', '2', '--batch_size', '4', '--n_sample', '4'], tmp,
                       timeout=1200)
    logs = glob.glob(os.path.join(tmp, 'Exp_*', '*_training_log.out'))
    assert len(logs) == 1, (out[-1000:], err[-2000:])
    text = open(logs[0]).read()
    lines = [ln for ln in text.splitlines() if ln.startswith('Iter #:')]
    assert len(lines) == 2, text[-1500:]
    for ln in lines:
        vals = [float(tok) for tok in ln.replace(':', ' ').split() if tok.replace('.', '', 1).replace('-', '', 1).isdigit()]
        assert all(np.isfinite(v) for v in vals), ln
    assert 'Total training time' in text
    assert glob.glob(os.path.join(tmp, 'Exp_*', 'sample', '000000.png'))
