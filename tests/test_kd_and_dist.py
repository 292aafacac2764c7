"""KD step parity (GPU, against oracle.kd_step) and the data-parallel host logic (CPU, gloo, world 2)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _small_nets(seed=0):
    import model
    torch.manual_seed(seed)
    teacher = model.Generator(32, 32, 2, generator_net_shape=[24] * 8)
    student = model.Generator(32, 32, 2, generator_net_shape=[13, 13, 13, 13, 9, 9, 6, 6])
    disc = model.Discriminator(32)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for net in (teacher, student):
            for n, p in net.named_parameters():
                if n.endswith('noise.weight') or n.endswith('activate.bias') or (n.endswith('.bias') and p.ndim == 4):
                    p.copy_(torch.randn(p.shape, generator=g) * 0.3)
        for n, p in disc.named_parameters():
            if n.endswith('bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    return teacher, student, disc


@pytest.mark.gpu
def test_kd_step_matches_oracle():
    from oracle import stylegan2_oracle as O
    from b200gan import config
    from b200gan.kd import KDStep
    teacher, student, disc = _small_nets()
    tp = {k: v.double() for k, v in teacher.state_dict().items()}
    sp = {k: v.double() for k, v in student.state_dict().items()}
    dp = {k: v.double() for k, v in disc.state_dict().items()}
    b = 4
    z = [torch.randn(b, 32, dtype=torch.float64), torch.randn(b, 32, dtype=torch.float64)]
    sn = [torch.randn(b, 1, n.shape[2], n.shape[3], dtype=torch.float64) for n in student.make_noise()]
    tn = [torch.randn(b, 1, n.shape[2], n.shape[3], dtype=torch.float64) for n in teacher.make_noise()]
    mask = (torch.rand(1, 1, 32, 32) > 0.3).double()
    loss_ref, gref = O.kd_step(sp, tp, dp, 32, z, sn, tn, 3, mask)

    dev = 'cuda'
    torch.backends.cudnn.allow_tf32 = False          # the discriminator's library convs in full fp32 for this check
    torch.backends.cuda.matmul.allow_tf32 = False
    kd = KDStep(student.to(dev), teacher.to(dev), disc.to(dev), mask=mask.float().to(dev))
    before = {n: p.detach().clone() for n, p in student.named_parameters()}
    f = lambda ts: [t.float().to(dev) for t in ts]
    with config.exact_fp32():
        loss = kd.step(f(z), 3, f(sn), f(tn))
    assert abs(float(loss) - float(loss_ref)) <= 1e-4 * abs(float(loss_ref)), (float(loss), float(loss_ref))
    worst = 0.0
    for n, p in student.named_parameters():
        if n not in gref:
            continue
        e = (p.grad.double().cpu() - gref[n]).abs().max() / max(gref[n].abs().max(), 1e-30)
        worst = max(worst, float(e))
        assert e <= 1e-3, f'{n}: {float(e):.3e}'
    # the fused Adam kernel applied the averaged gradient exactly like torch.optim.Adam would (step 1)
    lr, b1, b2 = 0.002 * 0.8, 0.0, 0.99 ** 0.8
    for n, p in student.named_parameters():
        g = p.grad
        m = (1 - b1) * g
        v = (1 - b2) * g * g
        exp = before[n] - (lr / (1 - b1)) * m / ((v.sqrt() / (1 - b2) ** 0.5) + 1e-8)
        assert torch.allclose(p.detach(), exp, rtol=1e-4, atol=1e-6), n
    print('worst KD gradient error', worst)


# ------------------------------------------------------------------ CPU / gloo, world_size 2
def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, os.path.join(ROOT, 'content-aware-gan-compression_b200'))
        os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                          LOCAL_RANK=str(rank))
        from b200gan import dist as D
        D.init_from_env(backend='gloo')
        assert D.get_rank() == rank and D.get_world_size() == world
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
        ref = [p.detach().clone() for p in net.parameters()]
        bucket = D.FlatBucket(net.parameters())
        for p, r in zip(net.parameters(), ref):          # re-homing keeps the values
            assert torch.equal(p.detach(), r)
        x = torch.arange(14, dtype=torch.float32).reshape(2, 7) * (rank + 1)
        bucket.zero_grad()
        net(x).square().sum().backward()
        local = bucket.flat_grad.clone()
        assert local.abs().sum() > 0                     # autograd accumulated straight into the bucket
        bucket.allreduce_mean_()
        gathered = [torch.zeros_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        assert torch.allclose(bucket.flat_grad, sum(gathered))
        # the KD step's protocol: gradients detached during backward, packed with one batched copy, then reduced
        bucket.detach_grads()
        net(x).square().sum().backward()
        assert all(p.grad is not None and p.grad.data_ptr() != bucket.flat_grad.data_ptr() for p in net.parameters())
        bucket.pack_grads()
        assert torch.equal(bucket.flat_grad, local)      # same local gradients as the accumulate-in-place form
        bucket.allreduce_mean_()
        assert torch.allclose(bucket.flat_grad, sum(gathered))
        assert next(net.parameters()).grad.data_ptr() == bucket.flat_grad.data_ptr()
        # gather_grad: average over ranks, one message
        for p in net.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        D.gather_grad(list(net.parameters()))
        for p in net.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, 1.5))
        # saliency sharding: whole batches per rank, merged in batch order
        mine = D.shard_batches(5)
        assert mine == list(range(rank, 5, world))
        local_scores = {i: [np.full(3, float(i))] for i in mine}
        merged = D.gather_scores_in_batch_order(local_scores, 5)
        assert [float(m[0][0]) for m in merged] == [0.0, 1.0, 2.0, 3.0, 4.0]
        red = D.reduce_loss_dict({'a': torch.tensor(float(rank)), 'b': torch.tensor(2.0)})
        if rank == 0:
            assert float(red['a']) == 0.5 and float(red['b']) == 2.0
        assert float(D.reduce_sum(torch.tensor(1.0))) == world
        D.synchronize()
        dist.destroy_process_group()
        q.put((rank, 'ok'))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))


def test_data_parallel_host_logic_gloo_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == 'ok', f'rank {rank}: {msg}'


def test_flat_bucket_single_process():
    from b200gan import dist as D
    net = torch.nn.Linear(3, 2)
    b = D.FlatBucket(net.parameters())
    assert b.numel == 8 + 4 and D.get_world_size() == 1 and b.allreduce_mean_() is None
    net(torch.ones(1, 3)).sum().backward()
    assert torch.equal(b.flat_grad[:6], torch.ones(6)) and torch.equal(b.flat_grad[8:10], torch.ones(2))
    b.zero_grad()
    assert float(b.flat_grad.abs().sum()) == 0.0 and net.weight.grad.data_ptr() == b.flat_grad.data_ptr()
    # detached protocol: backward hands gradients over, one batched copy packs them into the bucket
    b.detach_grads()
    assert net.weight.grad is None
    (2 * net(torch.ones(1, 3))).sum().backward()
    assert net.weight.grad.data_ptr() != b.flat_grad.data_ptr()
    b.pack_grads()
    assert torch.equal(b.flat_grad[:6], 2 * torch.ones(6)) and torch.equal(b.flat_grad[8:10], 2 * torch.ones(2))
    assert net.weight.grad.data_ptr() == b.flat_grad.data_ptr()
    b.detach_grads()
    net.weight.sum().backward()          # bias not reached: its segment must read zero after packing
    b.pack_grads()
    assert torch.equal(b.flat_grad[:6], torch.ones(6)) and float(b.flat_grad[8:10].abs().sum()) == 0.0


@pytest.mark.gpu
def test_saliency_pass_vs_oracle_and_sharding_invariance():
    """content_aware_scores == oracle.saliency_scores batch by batch; prune masks equal; the result does not
    depend on how whole batches are dealt to ranks."""
    import model
    from b200gan import saliency as S, dist as D
    from oracle import stylegan2_oracle as O
    torch.manual_seed(2)
    shape = [32, 32, 32, 32, 24, 24, 16, 16, 12, 12]
    gen = model.Generator(64, 64, 2, generator_net_shape=shape)
    with torch.no_grad():
        for n, p in gen.named_parameters():
            if n.endswith('noise.weight') or n.endswith('activate.bias') or (n.endswith('.bias') and p.ndim == 4):
                p.copy_(torch.randn_like(p) * 0.3)
    sd = {k: v.double() for k, v in gen.state_dict().items()}
    gen = gen.cuda()
    seed, n_sample, bs, prob = 11, 11, 3, 0.2           # batches [3, 3, 5]
    assert S.batch_sizes(n_sample, bs) == [3, 3, 5]
    per_batch = S.content_aware_scores(gen, n_sample, bs, prob, 'cuda', seed=seed, latent_dim=64)
    assert len(per_batch) == 3 and len(per_batch[0]) == len(shape)
    # oracle, same per-batch streams
    for idx, b in enumerate(S.batch_sizes(n_sample, bs)):
        rng = np.random.RandomState(seed + idx)      # the pass's per-batch stream: latents, noise maps, salt & pepper
        z = torch.from_numpy(rng.standard_normal((b, 64)).astype(np.float32))
        noise = [torch.from_numpy(rng.standard_normal((b, 1, n.shape[2], n.shape[3])).astype(np.float32))
                 for n in gen.make_noise()]

        def noisy_fn(img, rng=rng):
            return S.noisy_images(img, S.default_mask, prob, rng)
        ref = O.saliency_scores(sd, 64, z.double(), [n.double() for n in noise], noisy_fn)
        for l, (a, r) in enumerate(zip(per_batch[idx], ref)):
            err = np.abs(a - r).max() / np.abs(r).max()
            assert err <= 2e-4, f'batch {idx} layer {l}: {err:.2e}'
    tot = S.total_scores(per_batch)
    ref_tot = O.prune_mask_from_scores(tot, 0.7)
    for a, r in zip(S.prune_masks(tot, 0.7), ref_tot):
        assert np.array_equal(a, r)
    # sharding invariance: recompute the batches "rank 1 of 2" would own and merge
    assert D.shard_batches(3, rank=1, world=2) == [1]
    again = S.content_aware_scores(gen, n_sample, bs, prob, 'cuda', seed=seed, latent_dim=64)
    for b0, b1 in zip(per_batch, again):
        for a, r in zip(b0, b1):
            assert np.array_equal(a, r), 'saliency pass is not bit-reproducible'


def test_salt_pepper_rng_order_matches_reference_loop():
    """Vectorised salt/pepper maps consume numpy's stream exactly like the reference's pixel loop."""
    from b200gan import saliency as S
    mask = (np.arange(36).reshape(6, 6) % 3 != 0)
    np.random.seed(5)
    val, hit = S.salt_pepper_maps(mask, 0.4, np.random)
    np.random.seed(5)
    ref_val = np.random.randint(low=0, high=2, size=(6, 6)) * 2 - 1
    ref_hit = np.zeros((6, 6), dtype=bool)
    for h in range(6):
        for w in range(6):
            if mask[h, w] == True and (np.random.random() < 0.4):   # noqa: E712  (reference :166-168)
                ref_hit[h, w] = True
    assert np.array_equal(val, ref_val) and np.array_equal(hit, ref_hit)
