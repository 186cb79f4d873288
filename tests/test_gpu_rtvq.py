"""GPU parity of the K4 quantisers (through the reference-named API in svd_hybrid/rtvq.py and
quantization_utils.py) against the oracle: codes, scale and zero-point are compared BIT-EXACTLY
on identical inputs, NaN/inf edge cases included."""
import numpy as np
import pytest
import torch

from oracle import cref
from oracle import svd_hybrid_ref as R

pytestmark = pytest.mark.gpu


def _same_f32(a, b):
    a, b = np.float32(a), np.float32(b)
    return (np.isnan(a) and np.isnan(b)) or a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n", [1, 2, 3, 5, 7, 19, 100, 1023, 4097, 100003, 1 << 20])
@pytest.mark.parametrize("bits,stages", [(4, 2), (2, 3), (8, 1), (3, 4), (4, 3)])
def test_multistage_codes_bit_exact(cuda_device, n, bits, stages):
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    g = torch.Generator().manual_seed(n * 131 + bits * 7 + stages)
    x = torch.randn(n, generator=g) * 0.01
    ref = R.rtvq_quantize(x, bits, stages)
    new = rtvq.multistage_residual_quantization(x, bits, stages)
    assert len(new) == len(ref) == stages
    for a, b in zip(ref, new):
        assert b["stage"] == a["stage"]
        assert b["quantized"].dtype == torch.uint8 and b["quantized"].device.type == "cpu"
        assert b["scale"].ndim == 0 and b["zero_point"].ndim == 0
        assert torch.equal(a["quantized"], b["quantized"]), f"codes differ at stage {a['stage']}"
        assert _same_f32(a["scale"].item(), b["scale"].item())
        assert _same_f32(a["zero_point"].item(), b["zero_point"].item())
        if np.isfinite(a["residual_norm"]):
            assert abs(a["residual_norm"] - b["residual_norm"]) <= 1e-4 * max(abs(a["residual_norm"]), 1e-30)  # torch fp32 norm vs fp64 accumulation
    d_ref = R.rtvq_dequantize(ref)
    d_new = rtvq.multistage_residual_dequantization(new)
    assert torch.equal(torch.isnan(d_ref), torch.isnan(d_new))
    fin = torch.isfinite(d_ref)
    assert torch.equal(d_ref[fin], d_new[fin])
    # and the plain-C oracle agrees with both
    codes, sc, zp, rn, deq = cref.rtvq(x.numpy(), bits, stages)
    for s in range(stages):
        assert np.array_equal(codes[s], new[s]["quantized"].numpy().astype(np.int32))


def test_known_answer_1_to_5_at_4_bits(cuda_device):
    """tests/test_rtvq.py:35-46 of the reference: [1,2,3,4,5] @ 4 bit -> codes [0,4,7,11,15], scale 3.75, zp -4."""
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    q, s, z = rtvq.asymmetric_quantization(torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0]), 4)
    assert q.tolist() == [0, 4, 7, 11, 15] and s.item() == 3.75 and z.item() == -4.0
    assert q.dtype == torch.uint8 and s.ndim == 0 and z.ndim == 0
    d = rtvq.asymmetric_dequantization(q, s, z)
    assert (torch.tensor([1.0, 2.0, 3.0, 4.0, 5.0]) - d).abs().max() <= s.item() / 2 * 1.5


@pytest.mark.parametrize("x", [[0.5], [0.0, 0.0, 0.0], [1.0, 1.0], [float("nan"), 1.0, 2.0], [0.0, 1e-30]])
def test_degenerate_inputs_match_reference_nan_semantics(cuda_device, x):
    """constant / 1-element / NaN tensors: scale = inf, zero-point = +-inf or NaN, code 0 (rtvq.py:17-20)."""
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    t = torch.tensor(x)
    ref = R.rtvq_quantize(t, 4, 2)
    new = rtvq.multistage_residual_quantization(t, 4, 2)
    for a, b in zip(ref, new):
        assert torch.equal(a["quantized"], b["quantized"])
        assert _same_f32(a["scale"].item(), b["scale"].item()), (a["scale"], b["scale"])
        assert _same_f32(a["zero_point"].item(), b["zero_point"].item()), (a["zero_point"], b["zero_point"])
    d_ref, d_new = R.rtvq_dequantize(ref), rtvq.multistage_residual_dequantization(new)
    assert torch.equal(torch.isnan(d_ref), torch.isnan(d_new))


def test_empty_tensor(cuda_device):
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    assert rtvq.multistage_residual_quantization(torch.tensor([]), 4, 2) == []
    q = rtvq.RTVQQuantizer(4, 2)
    obj = q.quantize(torch.tensor([]))
    assert obj["payloads"] == [] and q.dequantize(obj).numel() == 0
    assert rtvq.estimate_compression_ratio(torch.tensor([]), obj) == 0


def test_quantizer_roundtrip_properties(cuda_device):
    """tests/test_rtvq.py:74-133 of the reference: more stages / more bits => lower error; rel err < 0.5."""
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    x = torch.randn(200, generator=torch.Generator().manual_seed(0))
    errs = []
    for st in (1, 2, 3):
        q = rtvq.RTVQQuantizer(4, st)
        errs.append((x - q.dequantize(q.quantize(x))).norm().item())
    assert errs[0] > errs[1] > errs[2] and errs[1] / x.norm().item() < 0.5
    e = {b: (x - rtvq.RTVQQuantizer(b, 1).dequantize(rtvq.RTVQQuantizer(b, 1).quantize(x))).norm().item() for b in (2, 4, 8)}
    assert e[8] < e[4] < e[2]
    obj = rtvq.RTVQQuantizer(4, 2).quantize(x.view(10, 20))
    assert obj["original_shape"] == torch.Size([10, 20]) and obj["original_dtype"] == "torch.float32"
    assert rtvq.RTVQQuantizer(4, 2).dequantize(obj).shape == (10, 20)
    assert abs(rtvq.estimate_compression_ratio(torch.zeros(5), {"num_bits": 4, "num_stages": 2}) - 20 / 21) < 1e-12


def test_device_residency_of_results(cuda_device):
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    x = torch.randn(1000, device="cuda")
    q, s, z = rtvq.asymmetric_quantization(x, 8)
    assert q.device.type == s.device.type == z.device.type == "cuda"
    ref_q, ref_s, ref_z = R.asym_quant(x.cpu(), 8)
    assert torch.equal(q.cpu(), ref_q) and _same_f32(s.item(), ref_s.item()) and _same_f32(z.item(), ref_z.item())


@pytest.mark.parametrize("bits", [8, 16, 4])
def test_root_quantization_utils(cuda_device, bits):
    """Root quantization_utils: absmax (int8/int16, no clamp, dequant multiplies) and asymmetric."""
    from svd_quantization_task_merging_b200 import quantization_utils as qu
    x = torch.randn(5000, generator=torch.Generator().manual_seed(bits)) * 3
    q_ref, s_ref = R.absmax_quant(x, bits)
    q, s = qu.absmax_quantization(x, bits)
    assert q.dtype == q_ref.dtype and torch.equal(q, q_ref) and _same_f32(s.item(), s_ref.item())
    assert torch.equal(qu.dequantize_absmax(q, s), R.absmax_dequant(q_ref, s_ref))
    qa_ref, sa_ref, za_ref = R.asym_quant(x, bits)
    qa, sa, za = qu.asymmetric_quantization(x, bits)
    assert qa.dtype == qa_ref.dtype and torch.equal(qa, qa_ref)
    assert _same_f32(sa.item(), sa_ref.item()) and _same_f32(za.item(), za_ref.item())
    assert torch.equal(qu.dequantize_asymmetric(qa, sa, za), R.asym_dequant(qa_ref, sa_ref, za_ref))


def test_full_size_tensor_roundtrip_property(cuda_device):
    """ViT-L-14's largest tensor (4,194,304 elements), 4-bit x 3 stages: size-independent properties --
    codes in range, every stage's residual norm shrinks, dequantised error bounded by the last step."""
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    x = torch.randn(4096 * 1024, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1)) * 0.01
    pay = rtvq.multistage_residual_quantization(x, 4, 3)
    norms = [p["residual_norm"] for p in pay]
    assert norms[0] > norms[1] > norms[2]
    assert all(int(p["quantized"].max()) <= 15 for p in pay)
    d = rtvq.multistage_residual_dequantization(pay, device="cuda")
    step = 1.0 / pay[-1]["scale"].item()
    assert (x - d).abs().max().item() <= 0.51 * step * 1.01


@pytest.mark.parametrize("bits,n", [(4, 1000), (2, 4099), (8, 77), (1, 130), (4, 1 << 18)])
def test_warp_packed_codes_equal_the_byte_codes(cuda_device, bits, n):
    """The optional bit-packed copy (assembled with warp shuffles) holds exactly the reference-layout codes."""
    from svd_quantization_task_merging_b200.svd_hybrid import rtvq
    x = torch.randn(n, generator=torch.Generator().manual_seed(n + bits)) * 0.01
    pay = rtvq.multistage_residual_quantization(x, bits, 2, packed=True)
    for p in pay:
        words = p["packed"].numpy().view(np.uint32)
        codes = p["quantized"].numpy().astype(np.uint32)
        unpacked = np.zeros(len(words) * (32 // bits), np.uint32)
        for i in range(32 // bits):
            unpacked[i:: 32 // bits] = (words >> (i * bits)) & ((1 << bits) - 1)
        assert np.array_equal(unpacked[:n], codes)
        assert (unpacked[n:] == 0).all()
