"""Root-level TaskVector / quantised containers (reference tests/test_task_vectors.py): exact arithmetic,
8-bit whole-tensor round trips, dtype skipping, nn.Module / file inputs -- and bit-equality of every
quantised payload with the oracle's restatement of quantization_utils."""
import pytest
import torch

from oracle import svd_hybrid_ref as R

pytestmark = pytest.mark.gpu


def _states():
    g = torch.Generator().manual_seed(0)
    pre = {"w": torch.randn(64, 48, generator=g), "b": torch.randn(48, generator=g),
           "steps": torch.tensor([7], dtype=torch.int64), "flags": torch.zeros(4, dtype=torch.uint8)}
    fin = {k: (v + 0.01 * torch.randn(v.shape, generator=g) if v.is_floating_point() else v.clone()) for k, v in pre.items()}
    return pre, fin


def test_task_vector_arithmetic_is_exact(cuda_device):
    from task_vectors import TaskVector
    pre, fin = _states()
    tv = TaskVector(pre, fin, task_name="A", verbose=False)
    assert set(tv.vector) == {"w", "b"}                                     # int64 / uint8 skipped
    assert torch.equal(tv.vector["w"], fin["w"] - pre["w"])
    two = tv + tv
    assert two.task_name == "A+A" and torch.equal(two.vector["w"], tv.vector["w"] * 2)
    assert all(torch.equal(v, torch.zeros_like(v)) for v in (tv - tv).vector.values())
    assert torch.equal((0.5 * tv).vector["b"], tv.vector["b"] * 0.5) and torch.equal((tv * 0.5).vector["b"], tv.vector["b"] * 0.5)
    out = tv.apply_to(pre)
    assert torch.equal(out["w"], pre["w"] + tv.vector["w"]) and torch.equal(out["steps"], pre["steps"])
    assert TaskVector(pre, fin, skip_int64=False, skip_uint8=False, verbose=False).vector.keys() == pre.keys()


def test_task_vector_from_files_and_modules(cuda_device, tmp_path):
    from task_vectors import TaskVector
    m0, m1 = torch.nn.Linear(8, 4), torch.nn.Linear(8, 4)
    torch.save(m0.state_dict(), tmp_path / "pre.pt")
    torch.save({"state_dict": m1.state_dict()}, tmp_path / "fin.pt")
    a = TaskVector(str(tmp_path / "pre.pt"), str(tmp_path / "fin.pt"), verbose=False)
    b = TaskVector(m0, m1, verbose=False)
    for k in a.vector:
        assert torch.equal(a.vector[k], b.vector[k])
    merged = a.apply_to(m0)
    assert torch.allclose(merged["weight"], m1.state_dict()["weight"], atol=1e-6)


@pytest.mark.parametrize("method", ["asymmetric", "absmax"])
def test_quantized_containers_match_oracle_bit_for_bit(cuda_device, method):
    from task_vectors import QuantizedBaseAndTaskVector, QuantizedFinetunedModel, QuantizedTaskVector, TaskVector
    pre, fin = _states()
    qf = QuantizedFinetunedModel(fin, qbit=8, method=method)
    assert set(qf.quantized_weights) == {"w", "b"}
    for k, pay in qf.quantized_weights.items():
        if method == "asymmetric":
            q, s, z = R.asym_quant(fin[k], 8)
            assert torch.equal(pay["quantized"], q) and pay["scale"].item() == s.item() and pay["zero_point"].item() == z.item()
        else:
            q, s = R.absmax_quant(fin[k], 8)
            assert torch.equal(pay["quantized"], q) and pay["scale"].item() == s.item()
        assert pay["shape"] == fin[k].shape
    deq = qf.dequantize()
    if method == "asymmetric":                     # the reference's absmax dequantiser multiplies (sic): no round trip
        assert ((deq["w"] - fin["w"]).norm() / fin["w"].norm()).item() < 0.1
        tv = qf.get_task_vector(pre)
        assert ((tv["w"] - (fin["w"] - pre["w"])).norm() / fin["w"].norm()).item() < 0.1
    tvec = TaskVector(pre, fin, verbose=False)
    qb = QuantizedBaseAndTaskVector(pre, tvec, base_qbit=8, task_qbit=8, method=method)
    rec = qb.dequantize()
    assert set(rec) == {"w", "b"}
    if method == "asymmetric":
        assert ((rec["w"] - fin["w"]).norm() / fin["w"].norm()).item() < 0.2
        qt = QuantizedTaskVector({k: {kk: vv for kk, vv in p.items() if kk != "shape"} for k, p in qb.quantized_task.items()})
        out = qt.apply_to(pre)
        assert ((out["w"] - fin["w"]).norm() / fin["w"].norm()).item() < 0.1 and torch.equal(out["steps"], pre["steps"])
