"""Data-parallel bookkeeping on CPU: world_size 2 (and 3) over gloo.

The flat grad buffer (torchlsq.dp.FlatGradBuffer) is device-agnostic torch code; here each rank
fills its slices with the ORACLE's gradients of its batch shard (on a GPU the sm_100a kernels
write the same slices directly) and the all-reduced buffer must equal sum_r oracle(shard_r)
(every rank scales with its LOCAL numel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _site_grads(x, g, scale, shift, per_channel, cfg_kw, x2=None, relu=False):
    """x2 / relu: the site sits behind a fused prologue (residual add and / or ReLU) - sharding by batch and the
    flat-buffer all-reduce are the same, the per-rank gradients are those of fake_quant(relu(x + x2))."""
    from oracle import lsq_oracle as O
    if per_channel:
        outer, C, inner = 1, x.shape[0], int(np.prod(x.shape[1:]))
    else:
        outer, C, inner = 1, 1, x.size
    if x2 is not None:
        _, gs, gb = O.backward_add(g.reshape(-1), x.reshape(-1), x2.reshape(-1), scale, shift, O.cfg(**cfg_kw), outer, C, inner,
                                   per_channel, with_relu=relu)
    elif relu:
        _, gs, gb = O.backward_relu(g.reshape(-1), x.reshape(-1), scale, shift, O.cfg(**cfg_kw), outer, C, inner, per_channel)
    else:
        _, gs, gb = O.backward(g.reshape(-1), x.reshape(-1), scale, shift, O.cfg(**cfg_kw), outer, C, inner, per_channel)
    return gs, gb


def _prologue_of(i, x, lo, hi, fused):
    """fused runs: site 0 is relu -> fq, site 1 the residual join relu(x + x2) -> fq (x2 derived from x so every process agrees)."""
    if not fused:
        return {}
    if i == 0:
        return dict(relu=True)
    return dict(x2=np.ascontiguousarray(np.flip(x, axis=1) * 0.5)[lo:hi], relu=True)


def _make_problem():
    rng = np.random.default_rng(0)
    acts = [(rng.standard_normal((12, 8, 5, 5)).astype(np.float32), rng.standard_normal((12, 8, 5, 5)).astype(np.float32)),
            (rng.standard_normal((12, 40)).astype(np.float32), rng.standard_normal((12, 40)).astype(np.float32))]
    w = (rng.standard_normal((6, 4, 3, 3)) * 0.05).astype(np.float32)
    gw_per_rank = [rng.standard_normal((6, 4, 3, 3)).astype(np.float32) for _ in range(3)]
    return acts, w, gw_per_rank


ACT_CFG = dict(quant_min=0, quant_max=127, type_min=0, type_max=255)
W_CFG = dict(quant_min=-128, quant_max=127, type_min=-128, type_max=127, sym=True)


def _worker(rank, world, port, average, out_dir, fused=False):
    import sys
    for p in (str(PKG), str(ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from torchlsq.dp import FlatGradBuffer, shard_batch
    acts, w, gw = _make_problem()
    flat = FlatGradBuffer([("act0", 1), ("act1", 1), ("w0", 6)], "cpu")
    for i, (x, g) in enumerate(acts):
        lo, hi = shard_batch(x.shape[0], rank, world)
        gs, gb = _site_grads(x[lo:hi], g[lo:hi], [0.05], [-0.7], False, ACT_CFG, **_prologue_of(i, x, lo, hi, fused))
        s_view, b_view = flat.views(f"act{i}")
        s_view.copy_(torch.from_numpy(gs).float())
        b_view.copy_(torch.from_numpy(gb).float())
    gs, gb = _site_grads(w, gw[rank], [0.002] * 6, [0.0] * 6, True, W_CFG)     # weights replicated, per-rank upstream grads
    flat.gscale("w0").copy_(torch.from_numpy(gs).float())
    flat.gshift("w0").copy_(torch.from_numpy(gb).float())
    h = flat.all_reduce(average=average, side_stream=True)
    h.wait()
    torch.save(flat.flat.clone(), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,average,fused", [(2, False, False), (2, True, False), (3, False, False), (2, False, True)])
def test_flat_grad_allreduce_equals_sum_of_shard_grads(tmp_path, world, average, fused):
    import sys
    for p in (str(PKG), str(ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from torchlsq.dp import expected_allreduced, shard_batch
    port = _free_port()
    mp.spawn(_worker, args=(world, port, average, str(tmp_path), fused), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    for r in range(1, world):
        assert torch.equal(got[0], got[r])                    # every rank holds the same reduced buffer
    acts, w, gw = _make_problem()
    per_rank = []
    for r in range(world):
        vals = []
        for i, (x, g) in enumerate(acts):
            lo, hi = shard_batch(x.shape[0], r, world)
            gs, gb = _site_grads(x[lo:hi], g[lo:hi], [0.05], [-0.7], False, ACT_CFG, **_prologue_of(i, x, lo, hi, fused))
            vals += [gs, gb]
        gs, gb = _site_grads(w, gw[r], [0.002] * 6, [0.0] * 6, True, W_CFG)
        vals += [gs, gb]
        per_rank.append(torch.from_numpy(np.concatenate(vals)).float())
    want = expected_allreduced(per_rank, average)
    assert torch.allclose(got[0].double(), want, rtol=1e-6, atol=1e-9)
    # local-numel scaling: the sharded sum is NOT the unsharded gradient (it is larger by ~sqrt(world))
    x, g = acts[0]
    full, _ = _site_grads(x, g, [0.05], [-0.7], False, ACT_CFG, **_prologue_of(0, x, 0, x.shape[0], fused))
    ratio = float(want[0]) * (world if average else 1) / float(full[0])
    assert abs(ratio - np.sqrt(world)) < 0.2 * np.sqrt(world)


def test_shard_batch_covers_everything():
    import sys
    sys.path.insert(0, str(PKG))
    from torchlsq.dp import shard_batch
    for n in (1, 7, 256, 2048, 2050):
        for world in (1, 2, 3, 8):
            spans = [shard_batch(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_flat_buffer_layout_and_param_grads():
    import sys
    sys.path.insert(0, str(PKG))
    from torchlsq.dp import FlatGradBuffer
    flat = FlatGradBuffer([("a", 1), ("w", 4), ("b", 1)], "cpu")
    assert flat.numel == 12 and flat.offsets["w"] == (2, 4)
    flat.gscale("w").fill_(2.0)
    flat.gshift("b").fill_(5.0)
    assert flat.flat.tolist() == [0, 0, 2, 2, 2, 2, 0, 0, 0, 0, 0, 5]
    s, b = torch.nn.Parameter(torch.ones(4)), torch.nn.Parameter(torch.zeros(4))
    flat.scatter_to_params({"w": (s, b)})
    assert s.grad.data_ptr() == flat.gscale("w").data_ptr() and s.grad.tolist() == [2, 2, 2, 2]
    with pytest.raises(ValueError):
        FlatGradBuffer([("a", 1), ("a", 2)], "cpu")
    assert flat.all_reduce().wait() is None                 # no process group: a no-op handle


def test_allreduce_bound_holds_where_the_relative_check_cannot():
    """The round-1 bench check `max |got - want| / |want| < 1e-6` killed the N >= 4 runs: an fp32 sum of W >= 3 addends has no
    error bound relative to the RESULT once shards cancel.  `torchlsq.dp.allreduce_bound` is the bound that is valid for any
    summation order; simulated here with fp32 tree / ring / sequential sums of buffers shaped like the bench's (55 262 floats)."""
    import sys
    sys.path.insert(0, str(PKG))
    from torchlsq.dp import allreduce_bound
    gen = torch.Generator().manual_seed(0)
    for world in (2, 4, 8):
        shards = [torch.randn(55262, generator=gen) * torch.rand(55262, generator=gen) for _ in range(world)]
        want, bound = allreduce_bound(shards)
        orders = []
        seq = shards[0].clone()
        for t in shards[1:]:
            seq = seq + t                                     # sequential (ring-like)
        orders.append(seq)
        tree = list(shards)
        while len(tree) > 1:                                  # pairwise tree
            tree = [tree[i] + tree[i + 1] if i + 1 < len(tree) else tree[i] for i in range(0, len(tree), 2)]
        orders.append(tree[0])
        rev = shards[-1].clone()
        for t in reversed(shards[:-1]):
            rev = rev + t
        orders.append(rev)
        worst_rel = 0.0
        for got in orders:
            err = (got.double() - want).abs()
            assert bool((err <= bound).all()), world
            worst_rel = max(worst_rel, float((err / (want.abs() + 1e-12)).max()))
        if world == 2:
            assert worst_rel <= 2.0 ** -24 * 1.0001            # one correctly rounded addition
        else:
            assert worst_rel > 1e-6, (world, worst_rel)        # the old check fails by construction
