"""GPU parity of the integer-export kernels (lsqb200_quantize / _dequantize / _qparams, include/lsq_b200.h)
through the C ABI.  Integer work: every comparison is bit-exact.

  * sem 'lsq'       vs the oracle (pinned to the reference CPU op's forward, tests/test_oracle_export.py) and through
                    the round-trip property dequantize(quantize(x)) == torchlsq.functional.lsq(x) at full size;
  * sem 'torch'     vs torch's own CUDA quantize_per_tensor / quantize_per_channel on the reference module's qparams
                    (what torch.quantization.convert does) - finite inputs;
  * sem 'torch_cpu' vs the committed golden codes made by torch's CPU quantizers (tests/golden/ref_export.npz);
  * qparams         vs the reference module's calculate_qparams() golden values.
"""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import lsq_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLDEN / "ref_export.npz")


def _ex():
    from torchlsq import export
    return export


CASES = [f"t{i}" for i in range(6)] + ["c", "a"]


def _case(gold, name):
    x = torch.from_numpy(gold[f"tq_{name}/x"]).to(DEV)
    s = torch.from_numpy(gold[f"qp_{name}/scale_in"]).to(DEV)
    b = torch.from_numpy(gold[f"qp_{name}/shift_in"]).to(DEV)
    if name == "c":
        return x, s, b, dict(quant_min=-128, quant_max=127, axis=0, is_perchannel=True), (1, x.shape[0], x.shape[1])
    if name == "a":
        return x, s, b, dict(quant_min=0, quant_max=255, axis=1, is_perchannel=True), tuple(x.shape)
    return x, s, b, dict(quant_min=0, quant_max=255), (1, 1, x.numel())


@pytest.mark.parametrize("name", CASES)
def test_qparams_device_equals_reference_module(gold, name):
    x, s, b, kw, _ = _case(gold, name)
    tmin, tmax = (-128, 127) if name == "c" else (0, 255)
    so, zp = _ex().qparams(s, b, tmin, tmax)
    assert np.array_equal(so.cpu().numpy().view(np.uint32), gold[f"qp_{name}/scale"].view(np.uint32))
    assert np.array_equal(zp.cpu().numpy(), gold[f"qp_{name}/zp"])


@pytest.mark.parametrize("name", CASES)
def test_torch_cpu_semantics_equals_golden_codes(gold, name):
    x, s, b, kw, _ = _case(gold, name)
    codes = _ex().quantize(x, s, b, semantics='torch_cpu', **kw)
    want = gold[f"tq_{name}/codes"]
    assert codes.dtype == (torch.int8 if name == "c" else torch.uint8)
    assert np.array_equal(codes.cpu().numpy(), want), np.flatnonzero(codes.cpu().numpy().reshape(-1) != want.reshape(-1))[:8]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("sem", ["lsq", "torch", "torch_cpu"])
def test_codes_equal_oracle(gold, name, sem):
    x, s, b, kw, (outer, C, inner) = _case(gold, name)
    if sem == "torch":
        x = torch.nan_to_num(x, nan=0.0, posinf=3e38, neginf=-3e38)   # torch's CUDA cast of NaN / inf is not pinned
    codes = _ex().quantize(x, s, b, semantics=sem, **kw).cpu().numpy().astype(np.int32).reshape(-1)
    c = O.cfg(kw["quant_min"], kw["quant_max"])
    want = O.quantize(x.cpu().numpy().reshape(-1), s.cpu().numpy(), b.cpu().numpy(), c, outer, C, inner,
                      kw.get("is_perchannel", False), sem={"lsq": O.SEM_LSQ, "torch": O.SEM_TORCH_CUDA, "torch_cpu": O.SEM_TORCH_CPU}[sem])
    assert np.array_equal(codes, want), np.flatnonzero(codes != want)[:8]


@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("n", [1, 31, 4099, 1 << 20])
@pytest.mark.parametrize("offset", [0, 1, 3])
def test_torch_cuda_semantics_equals_torch_quantize_per_tensor(dtype, n, offset):
    """The fp32-only reconstruction of nearbyint(double(x) / double(scale)) must agree with torch's CUDA kernel on every
    element, ties included (x = (k + 0.5) * scale planted)."""
    ex = _ex()
    gen = torch.Generator().manual_seed(n + offset)
    for sv, bv in ((0.03, -1.7), (0.25, -0.6), (0.0173, 2.2), (1.0 / 3.0, -40.0)):
        s = torch.tensor([sv], device=DEV)
        b = torch.tensor([bv], device=DEV)
        base = (torch.randn(n + offset, generator=gen) * 3.0).to(DEV)
        k = torch.arange(n + offset, device=DEV) % 200
        base[::3] = ((k[::3].float() - 60.0 + 0.5) * s).float()             # exact and near ties
        x = base.to(dtype)[offset:]                                         # misaligned views
        so, zp = ex.qparams(s, b, 0, 255)
        ref = torch.quantize_per_tensor(x.float(), float(so[0]), int(zp[0]), torch.quint8).int_repr()
        codes = ex.quantize(x, s, b, 0, 255, semantics='torch')
        assert torch.equal(codes, ref), (sv, bv, (codes != ref).nonzero()[:4])
        y = ex.dequantize(codes, s, b, 0, 255, semantics='torch', dtype=torch.float32)
        yr = torch._make_per_tensor_quantized_tensor(codes, float(so[0]), int(zp[0])).dequantize()
        assert torch.equal(y, yr)


@pytest.mark.parametrize("shape,axis", [((64, 32, 3, 3), 0), ((8, 12, 7, 7), 1), ((5, 33), 0), ((4, 6, 10), 2)])
def test_torch_cuda_semantics_equals_torch_quantize_per_channel(shape, axis):
    ex = _ex()
    gen = torch.Generator().manual_seed(sum(shape))
    C = shape[axis]
    x = (torch.randn(*shape, generator=gen) * 0.4).to(DEV)
    s = (0.002 + 0.01 * torch.rand(C, generator=gen)).to(DEV)
    b = torch.zeros(C, device=DEV)
    so, zp = ex.qparams(s, b, -128, 127)
    ref = torch.quantize_per_channel(x, so.double(), zp, axis, torch.qint8).int_repr()
    codes = ex.quantize(x, s, b, -128, 127, axis=axis, is_perchannel=True, semantics='torch')
    assert codes.dtype == torch.int8 and torch.equal(codes, ref)
    q = ex.to_quantized_tensor(x, _FakeFQ(s, b, axis))
    assert q.is_quantized and torch.equal(q.int_repr(), ref)
    assert torch.equal(q.dequantize(), ex.dequantize(codes, s, b, -128, 127, axis=axis, is_perchannel=True, semantics='torch'))


class _FakeFQ:
    """just the attributes export.to_quantized_tensor reads from an LSQFakeQuantizer"""
    dtype = torch.qint8
    quant_min, quant_max = -128, 127
    is_perchannel = True

    def __init__(self, s, b, axis):
        self.scale, self.shift, self.ch_axis = s, b, axis


@pytest.mark.parametrize("dtype,pdtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.float32),
                                          (torch.float16, torch.float32), (torch.float16, torch.float16)])
def test_lsq_roundtrip_equals_fake_quant_forward_full_size(dtype, pdtype):
    """Size-independent property at BASELINE configs[0] size: dequantize(quantize(x)) is the training forward, bit for bit
    (so the exported integers are exactly the ones the model was trained against)."""
    from torchlsq.functional import lsq
    ex = _ex()
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(32, 64, 56, 56, generator=gen).to(dtype).to(DEV)
    s = torch.tensor([0.03], dtype=pdtype, device=DEV)
    b = torch.tensor([-1.7], dtype=pdtype, device=DEV)
    codes = ex.quantize(x, s, b, 0, 127, 0, 255)
    assert codes.dtype == torch.uint8 and int(codes.max()) <= 127
    y = ex.dequantize(codes, s, b, 0, 127, 0, 255, dtype=dtype)
    assert torch.equal(y, lsq(x, s, b, 0, 127, 0, 255))
    # per-channel, channels-last memory: codes keep x's strides
    xc = x[:4].to(memory_format=torch.channels_last)
    sc = (0.02 + 0.02 * torch.rand(64, generator=gen)).to(pdtype).to(DEV)
    bc = (-torch.rand(64, generator=gen)).to(pdtype).to(DEV)
    cc = ex.quantize(xc, sc, bc, 0, 127, 0, 255, axis=1, is_perchannel=True)
    assert cc.stride() == xc.stride()
    yc = ex.dequantize(cc, sc, bc, 0, 127, 0, 255, axis=1, is_perchannel=True, dtype=dtype)
    assert torch.equal(yc, lsq(xc, sc, bc, 0, 127, 0, 255, axis=1, is_perchannel=True))


def test_export_errors():
    ex = _ex()
    x = torch.randn(16, device=DEV)
    s = torch.tensor([0.1], device=DEV)
    b = torch.tensor([0.0], device=DEV)
    with pytest.raises(RuntimeError):
        ex.quantize(x.cpu(), s, b)                              # no CPU path
    with pytest.raises(RuntimeError):
        ex.quantize(x, s, b, 0, 511)                            # does not fit 8 bits
    with pytest.raises(RuntimeError):
        ex.quantize(x, s, b, -128, 127, code_dtype=torch.uint8)  # signed range into unsigned codes
    with pytest.raises(RuntimeError):
        ex.quantize(x.half(), s.half(), b.half(), semantics='torch')  # torch semantics need fp32 qparams
    assert ex.quantize(x[:0], s, b).numel() == 0
