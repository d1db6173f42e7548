"""Data-parallel parity on the REAL path: sm_100a kernels write grad_scale / grad_shift straight into the flat buffer
(torchlsq.dp.FlatGradBuffer), NCCL all-reduces it, and every rank's result must equal sum_r oracle(shard_r) - each rank scaled
with its LOCAL numel, as the reference op would be under DDP (/root/reference/torchlsq/csrc/ops/cuda/lsq_cuda.cu:124,274;
SURVEY.md 7.2-7).  Needs >= 2 GPUs (skipped otherwise); world 4 runs where 4 GPUs are visible.

Bound per element (fp32 NCCL sum of W addends, any order):  |got - want| <= 1e-6 |want| + W 2^-24 sum_r |oracle(shard_r)|,
plus the kernels' own 1e-6 relative agreement with the oracle per rank (checked separately, before the reduction)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu

ACT = dict(quant_min=0, quant_max=127, type_min=0, type_max=255)
WGT = dict(quant_min=-128, quant_max=127, type_min=-128, type_max=127, sym=True)
BATCH, ACT_SHAPES, W_SHAPES = 24, [(16, 14, 14), (40,), (8, 7, 7)], [(12, 8, 3, 3), (10, 40)]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(5)
    acts = [(np.maximum(rng.standard_normal((BATCH,) + s), 0).astype(np.float32), rng.standard_normal((BATCH,) + s).astype(np.float32))
            for s in ACT_SHAPES]
    wts = [(rng.standard_normal(s) * 0.05).astype(np.float32) for s in W_SHAPES]
    gws = [[rng.standard_normal(s).astype(np.float32) for s in W_SHAPES] for _ in range(8)]      # per-rank upstream weight grads
    return acts, wts, gws


def _oracle_rank(rank, world, dtype_name):
    """Flat buffer a rank must produce, from the oracle on that rank's shard (bf16 activations: the same bits the GPU sees)."""
    import sys
    for p in (str(PKG), str(ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import lsq_oracle as O
    from torchlsq.dp import shard_batch
    acts, wts, gws = _problem()
    vals, mags = [], []
    for x, g in acts:
        lo, hi = shard_batch(BATCH, rank, world)
        xs, gs_ = torch.from_numpy(x[lo:hi]), torch.from_numpy(g[lo:hi])
        if dtype_name == "bf16":
            xs, gs_ = xs.bfloat16(), gs_.bfloat16()
        xb, dt = O.to_bits(xs)
        gb, _ = O.to_bits(gs_)
        _, s, b, ms, mb = O.backward(gb.reshape(-1), xb.reshape(-1), [0.05], [-0.7], O.cfg(**ACT), dt=dt, with_abs=True)
        vals += [s, b]
        mags += [ms, mb]
    for w, gw in zip(wts, gws[rank]):
        C = w.shape[0]
        _, s, b, ms, mb = O.backward(gw.reshape(-1), w.reshape(-1), [0.002] * C, [0.0] * C, O.cfg(**WGT), 1, C, w.size // C, True,
                                     with_abs=True)
        vals += [s, b]
        mags += [ms, mb]
    cat = lambda vs: np.concatenate([np.asarray(v, dtype=np.float64).reshape(-1) for v in vs])
    return cat(vals), cat(mags)


def _worker(rank, world, port, dtype_name, out_dir):
    import sys
    for p in (str(PKG), str(ROOT)):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from torchlsq import _cabi
    from torchlsq.dp import FlatGradBuffer, shard_batch
    from torchlsq.multi import LSQPlan, Site
    lib = _cabi.load()
    acts, wts, gws = _problem()
    adt = torch.bfloat16 if dtype_name == "bf16" else torch.float32
    flat = FlatGradBuffer([(f"a{i}", 1) for i in range(len(acts))] + [(f"w{i}", w.shape[0]) for i, w in enumerate(wts)], dev)
    ws = torch.zeros(lib.lsqb200_workspace_bytes(), dtype=torch.uint8, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    q = _cabi.qargs(0, 127, 0, 255, True, 1.0, False, False, False)
    s_act, b_act = torch.tensor([0.05], device=dev), torch.tensor([-0.7], device=dev)
    keep = []
    for i, (x, g) in enumerate(acts):                    # activations: sharded by batch, per-site C-ABI calls into the flat slices
        lo, hi = shard_batch(BATCH, rank, world)
        xd, gd = torch.from_numpy(x[lo:hi]).to(adt).to(dev).contiguous(), torch.from_numpy(g[lo:hi]).to(adt).to(dev).contiguous()
        gx = torch.empty_like(xd)
        gs, gb = flat.views(f"a{i}")
        rc = lib.lsqb200_bwd_tensor(gd.data_ptr(), xd.data_ptr(), gx.data_ptr(), s_act.data_ptr(), b_act.data_ptr(), gs.data_ptr(),
                                    gb.data_ptr(), xd.numel(), _cabi.BF16 if dtype_name == "bf16" else _cabi.F32, _cabi.F32, q,
                                    ws.data_ptr(), ws.numel(), sp)
        _cabi.check(rc, "bwd_tensor")
        keep += [xd, gd, gx]
    sites = []
    for i, (w, gw) in enumerate(zip(wts, gws[rank])):    # weights: replicated, one multi-tensor plan, per-rank upstream grads
        wd, gd = torch.from_numpy(w).to(dev), torch.from_numpy(gw).to(dev)
        gs, gb = flat.views(f"w{i}")
        sites.append(Site(x=wd, y=torch.empty_like(wd), grad=gd, gx=torch.empty_like(wd), scale=torch.full((w.shape[0],), 0.002, device=dev),
                          shift=torch.zeros(w.shape[0], device=dev), gscale=gs, gshift=gb, quant_min=-128, quant_max=127, type_min=-128,
                          type_max=127, axis=0, is_affine=False, is_perchannel=True))
    plan = LSQPlan(sites)
    plan.backward()
    torch.cuda.synchronize()
    local = flat.flat.clone()
    flat.all_reduce(side_stream=True).wait()
    torch.cuda.synchronize()
    torch.save({"local": local.cpu(), "reduced": flat.flat.cpu()}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("dtype_name", ["bf16", "f32"])
def test_nccl_allreduced_grads_equal_sum_of_oracle_shard_grads(tmp_path, world, dtype_name):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    mp.spawn(_worker, args=(world, _free_port(), dtype_name, str(tmp_path)), nprocs=world, join=True)
    got = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    both = [_oracle_rank(r, world, dtype_name) for r in range(world)]
    want_rank, mag_rank = [b[0] for b in both], [b[1] for b in both]
    for r in range(world):      # each rank's kernels against the oracle on that rank's shard, before any reduction
        loc = got[r]["local"].double().numpy()
        err = np.abs(loc - want_rank[r])
        # the single-GPU parity bar (tests/gpu_util.py::assert_grads_close): 1e-6 relative + the fp32 floor of the terms themselves
        assert np.all(err <= 1e-6 * np.abs(want_rank[r]) + 1e-7 * mag_rank[r] + 1e-30), (r, float(err.max()))
        assert torch.equal(got[r]["reduced"], got[0]["reduced"])                 # every rank holds the same reduced buffer
    want = np.sum(want_rank, axis=0)
    bound = 1e-6 * np.abs(want) + 1e-7 * np.sum(mag_rank, axis=0) + world * 2.0 ** -24 * np.sum(np.abs(want_rank), axis=0) + 1e-30
    err = np.abs(got[0]["reduced"].double().numpy() - want)
    assert np.all(err <= bound), float((err / bound).max())
    assert np.abs(want).max() > 0
