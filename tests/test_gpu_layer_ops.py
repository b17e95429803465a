"""CUDA glue kernels vs the oracle's torch restatement of the reference ops (same rounding points)."""
import pytest
import torch

from oracle import llama_ref as LR

pytestmark = pytest.mark.gpu


def _lib():
    from lookaheaddecoding_b200 import _cabi
    return _cabi.load(), _cabi.check


def _stream():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.parametrize("rows,hidden", [(1, 256), (120, 4096), (77, 5120)])
def test_rmsnorm_and_residual(rows, hidden):
    lib, check = _lib()
    torch.manual_seed(rows)
    x = torch.randn(rows, hidden, device="cuda").to(torch.bfloat16)
    d = (torch.randn(rows, hidden, device="cuda") * 0.5).to(torch.bfloat16)
    w = (1 + 0.1 * torch.randn(hidden, device="cuda")).to(torch.bfloat16)
    out = torch.empty_like(x)
    check(lib.lade_rmsnorm(_stream(), x.data_ptr(), 0, w.data_ptr(), 0, out.data_ptr(), rows, hidden, 1e-5))
    ref = LR.rms_norm(x, w, 1e-5)
    # fp32 reduction order differs from torch's; allow 1 bf16 ulp on a vanishing fraction
    assert (out != ref).float().mean() < 2e-3
    torch.testing.assert_close(out.float(), ref.float(), rtol=1e-2, atol=1e-3)
    # fused residual: h = bf16(x + d) written back, norm of h
    h = x.clone()
    check(lib.lade_rmsnorm(_stream(), h.data_ptr(), d.data_ptr(), w.data_ptr(), h.data_ptr(), out.data_ptr(),
                           rows, hidden, 1e-5))
    assert torch.equal(h, x + d)
    ref2 = LR.rms_norm(x + d, w, 1e-5)
    assert (out != ref2).float().mean() < 2e-3
    # gather variant
    idx = torch.tensor([rows - 1, 0, rows // 2], dtype=torch.int32, device="cuda")
    og = torch.empty(3, hidden, dtype=torch.bfloat16, device="cuda")
    check(lib.lade_rmsnorm_gather(_stream(), x.data_ptr(), d.data_ptr(), w.data_ptr(), idx.data_ptr(), og.data_ptr(),
                                  3, hidden, 1e-5))
    assert torch.equal(og, out[idx.long()])


@pytest.mark.parametrize("nh,nkv,D,dt", [(2, 2, 128, torch.bfloat16), (4, 2, 128, torch.bfloat16), (32, 32, 128, torch.bfloat16),
                                         (4, 2, 64, torch.bfloat16), (4, 2, 128, torch.float16), (4, 4, 64, torch.float16)])
def test_rope_append_bitexact(nh, nkv, D, dt):
    from lookaheaddecoding_b200 import _cabi
    lib, check = _lib()
    rows, q_pad, cap, max_pos, kv_len = 37, 40, 200, 512, 61
    torch.manual_seed(nh)
    fn = lib.lade_rope_append if dt == torch.bfloat16 else lib.lade_rope_append_f16
    qkv = torch.randn(rows, (nh + 2 * nkv) * D, device="cuda").to(dt)
    cos, sin = LR.rope_tables(D, max_pos, 10000.0, dt, "cuda")
    pos = torch.randint(0, max_pos, (rows,), dtype=torch.int32, device="cuda")
    meta = torch.zeros(_cabi.META_INTS, dtype=torch.int32, device="cuda")
    meta[_cabi.M_KV_LEN] = kv_len
    qo = torch.zeros(nh, q_pad, D, dtype=dt, device="cuda")
    kc = torch.zeros(nkv, cap, D, dtype=dt, device="cuda")
    vc = torch.zeros(nkv, cap, D, dtype=dt, device="cuda")
    check(fn(_stream(), qkv.data_ptr(), cos.data_ptr(), sin.data_ptr(), pos.data_ptr(),
                               meta.data_ptr(), qo.data_ptr(), kc.data_ptr(), vc.data_ptr(), rows, q_pad, nh, nkv, D,
                               cap, max_pos))
    q = qkv[:, : nh * D].view(rows, nh, D).transpose(0, 1)
    k = qkv[:, nh * D:(nh + nkv) * D].view(rows, nkv, D).transpose(0, 1)
    v = qkv[:, (nh + nkv) * D:].view(rows, nkv, D).transpose(0, 1)
    c, s = cos[pos.long()], sin[pos.long()]
    q_ref = (q * c) + (LR.rotate_half(q) * s)            # modeling_llama.py:344-345
    k_ref = (k * c) + (LR.rotate_half(k) * s)
    assert torch.equal(qo[:, :rows], q_ref)
    assert torch.equal(kc[:, kv_len:kv_len + rows], k_ref)
    assert torch.equal(vc[:, kv_len:kv_len + rows], v)
    assert kc[:, :kv_len].abs().sum() == 0 and kc[:, kv_len + rows:].abs().sum() == 0


def test_swiglu():
    lib, check = _lib()
    torch.manual_seed(3)
    rows, inter = 53, 11008
    gu = (torch.randn(rows, 2 * inter, device="cuda") * 2).to(torch.bfloat16)
    out = torch.empty(rows, inter, dtype=torch.bfloat16, device="cuda")
    check(lib.lade_swiglu(_stream(), gu.data_ptr(), out.data_ptr(), rows, inter))
    ref = torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]
    assert (out != ref).float().mean() < 1e-3          # expf vs torch's exp: <= 1 bf16 ulp, rarely
    torch.testing.assert_close(out.float(), ref.float(), rtol=1e-2, atol=1e-3)


def test_kv_compact():
    from lookaheaddecoding_b200 import _cabi
    lib, check = _lib()
    L, nkv, cap, D = 3, 2, 64, 128
    kv = torch.randn(L, 2, nkv, cap, D, device="cuda").to(torch.bfloat16)
    ref = kv.clone()
    res = torch.zeros(_cabi.RES_INTS, dtype=torch.int32, device="cuda")
    res[_cabi.R_MAX_HIT], res[_cabi.R_KV_SRC], res[_cabi.R_KV_DST] = 2, 40, 11
    check(lib.lade_kv_compact(_stream(), res.data_ptr(), kv[0, 0].data_ptr(), kv[0, 1].data_ptr(), kv.stride(0), L, nkv,
                              cap, D, 3))
    ref[:, :, :, 11:13] = ref[:, :, :, 40:42]
    assert torch.equal(kv, ref)
    res[_cabi.R_MAX_HIT] = 0
    check(lib.lade_kv_compact(_stream(), res.data_ptr(), kv[0, 0].data_ptr(), kv[0, 1].data_ptr(), kv.stride(0), L, nkv,
                              cap, D, 3))
    assert torch.equal(kv, ref)


def test_fp16_glue_kernels_match_the_restated_reference_ops():
    """The *_f16 entry points: same kernels instantiated on __half -- every rounding point in fp16."""
    lib, check = _lib()
    torch.manual_seed(3)
    rows, hidden, inter = 37, 1024, 1376
    x = torch.randn(rows, hidden, device="cuda").to(torch.float16)
    d = (torch.randn(rows, hidden, device="cuda") * 0.5).to(torch.float16)
    w = (1 + 0.1 * torch.randn(hidden, device="cuda")).to(torch.float16)
    out = torch.empty_like(x)
    h = x.clone()
    check(lib.lade_rmsnorm_f16(_stream(), h.data_ptr(), d.data_ptr(), w.data_ptr(), h.data_ptr(), out.data_ptr(), rows, hidden, 1e-5))
    assert torch.equal(h, x + d)
    ref = LR.rms_norm(x + d, w, 1e-5)
    assert ref.dtype == torch.float16 and (out != ref).float().mean() < 2e-3
    torch.testing.assert_close(out.float(), ref.float(), rtol=2e-3, atol=2e-3)
    idx = torch.tensor([rows - 1, 0, 5], dtype=torch.int32, device="cuda")
    og = torch.empty(3, hidden, dtype=torch.float16, device="cuda")
    check(lib.lade_rmsnorm_gather_f16(_stream(), x.data_ptr(), 0, w.data_ptr(), idx.data_ptr(), og.data_ptr(), 3, hidden, 1e-5))
    refg = LR.rms_norm(x[idx.long()], w, 1e-5)
    assert (og != refg).float().mean() < 2e-3
    gu = torch.randn(rows, 2 * inter, device="cuda").to(torch.float16)
    act = torch.empty(rows, inter, dtype=torch.float16, device="cuda")
    check(lib.lade_swiglu_f16(_stream(), gu.data_ptr(), act.data_ptr(), rows, inter))
    want = torch.nn.functional.silu(gu[:, :inter]) * gu[:, inter:]
    assert (act != want).float().mean() < 5e-3
    torch.testing.assert_close(act.float(), want.float(), rtol=2e-3, atol=2e-3)
    lg = (torch.randn(9, 1003, device="cuda") * 4).round().to(torch.float16)
    am = torch.zeros(9, dtype=torch.int32, device="cuda")
    check(lib.lade_argmax_rows_f16(_stream(), lg.data_ptr(), 9, 1000, 1003, am.data_ptr()))
    assert am.cpu().tolist() == torch.argmax(lg[:, :1000].float().cpu(), dim=-1).tolist()
