"""Kernel-level parity on the B200 (through the C ABI test hooks), each against a torch fp64/fp32
reference of the same op.  Tolerances are stated per precision mode:
  fp32    CUDA-core fp32 FMA            : <= 2e-6 of the output range
  bf16x3  tcgen05, hi/lo planes, 3 MMAs : <= 3e-5 of the output range (2^-16 per product)
  bf16    tcgen05, plain bf16           : <= 2e-2 of the output range
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.gpu_util import RawContext, rel_err

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-6, "bf16x3": 3e-5, "bf16": 2e-2}


@pytest.fixture(scope="module")
def raw():
    ctx = RawContext()
    yield ctx
    ctx.close()


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).numpy().astype(np.float32)


LINEAR_SHAPES = [  # (m, n, k) — encoder qkv / out / ffn, layer-0 K=560, narrow CTC FFN, ragged M and N tiles
    (300, 512, 512), (129, 1536, 560), (77, 128, 512), (260, 264, 128), (1, 512, 2048), (515, 2048, 512), (128, 256, 1024),
]


@pytest.fixture(params=["1cta", "2cta"])
def gemm_kernel(request, monkeypatch):
    """Force the single-CTA or the CTA-pair tcgen05 GEMM (the library picks by size otherwise)."""
    monkeypatch.setenv("FUNASR_B200_GEMM", request.param)
    return request.param


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "bf16"])
@pytest.mark.parametrize("m,n,k", LINEAR_SHAPES)
def test_linear_bias_relu_residual(raw, gemm_kernel, precision, m, n, k):
    if precision == "fp32" and gemm_kernel == "2cta":
        pytest.skip("fp32 mode does not use the tensor-core kernels")
    a, w = _rand((m, k), 1), _rand((n, k), 2, k ** -0.5)
    bias, resid = _rand((n,), 3), _rand((m, n), 4)
    ref = torch.from_numpy(a).double() @ torch.from_numpy(w).double().t() + torch.from_numpy(bias).double()
    got = raw.linear(a, w, bias, precision=precision)
    assert rel_err(got, ref) <= TOL[precision]
    got, planes = raw.linear(a, w, bias, resid=resid, relu=True, precision=precision, planes=True)
    ref2 = torch.relu(ref) + torch.from_numpy(resid).double()
    assert rel_err(got, ref2) <= TOL[precision]
    # the bf16 hi+lo planes written by the epilogue carry the same values to 16 mantissa bits
    assert np.abs(planes - got).max() <= 2.0 ** -15 * np.abs(got).max()


@pytest.mark.parametrize("m,n,k", [(1001, 512, 512), (1001, 1536, 560), (1001, 512, 2048), (2002, 2048, 512), (300, 264, 128)])
def test_linear_narrow_tiles_for_small_m_are_bit_identical_to_the_wide_tiles(raw, monkeypatch, m, n, k):
    """Calls of one to four segments run the single-CTA kernel with 64- or 128-column tiles (more SMs, shorter K loop per
    tile: the latency of such a call is the sum of ~310 dependent projections).  An output element sees the same
    sequence of tensor-core products whatever the tile width, so the three widths must agree bit for bit."""
    monkeypatch.setenv("FUNASR_B200_GEMM", "1cta")
    a, w = _rand((m, k), 41), _rand((n, k), 42, k ** -0.5)
    bias, resid = _rand((n,), 43), _rand((m, n), 44)
    ref = torch.relu(torch.from_numpy(a).double() @ torch.from_numpy(w).double().t() + torch.from_numpy(bias).double())
    ref = ref + torch.from_numpy(resid).double()
    outs = {}
    for tn in ("256", "128", "64"):
        monkeypatch.setenv("FUNASR_B200_GEMM_TN", tn)
        outs[tn] = raw.linear(a, w, bias, resid=resid, relu=True, precision="bf16x3", planes=True)
    monkeypatch.delenv("FUNASR_B200_GEMM_TN")
    outs["auto"] = raw.linear(a, w, bias, resid=resid, relu=True, precision="bf16x3", planes=True)
    assert rel_err(outs["256"][0], ref) <= TOL["bf16x3"]
    for tn in ("128", "64", "auto"):
        assert np.array_equal(outs[tn][0], outs["256"][0]) and np.array_equal(outs[tn][1], outs["256"][1])


@pytest.mark.parametrize("m,n,k", [(10240, 512, 512), (9999, 1536, 512), (20000, 512, 2048)])
def test_linear_pair_kernel_full_waves_and_split_tail(raw, monkeypatch, m, n, k):
    """Sizes at which the library itself picks the CTA-pair kernel: several waves of 256 x 256 tiles and a
    last partial wave cut into 128-column half tiles (80, 240 and 158 pair tiles on 74 pairs)."""
    monkeypatch.delenv("FUNASR_B200_GEMM", raising=False)
    a, w = _rand((m, k), 31), _rand((n, k), 32, k ** -0.5)
    bias, resid = _rand((n,), 33), _rand((m, n), 34)
    ref = torch.relu(torch.from_numpy(a).double() @ torch.from_numpy(w).double().t() + torch.from_numpy(bias).double())
    ref = ref + torch.from_numpy(resid).double()
    got, planes = raw.linear(a, w, bias, resid=resid, relu=True, precision="bf16x3", planes=True)
    assert rel_err(got, ref) <= TOL["bf16x3"]
    assert np.abs(planes - got).max() <= 2.0 ** -15 * np.abs(got).max()


@pytest.fixture(params=["rescore", "full"])
def vocab_mode(request, monkeypatch):
    """bf16x3 vocabulary argmax: one-product pass + exact rescoring of the candidate columns (default), or the
    three-product projection with the fused running argmax (FUNASR_B200_VOCAB=full)."""
    if request.param == "full":
        monkeypatch.setenv("FUNASR_B200_VOCAB", "full")
    else:
        monkeypatch.delenv("FUNASR_B200_VOCAB", raising=False)
    return request.param


def test_vocab_argmax_pair_kernel_split_tail(raw, monkeypatch, vocab_mode):
    monkeypatch.delenv("FUNASR_B200_GEMM", raising=False)
    m, n = 2000, 5037                      # 8 x 20 pair tiles: two full waves of 74 and 12 tiles cut in halves
    a, w, bias = _rand((m, 512), 35), _rand((n, 512), 36, 512 ** -0.5), _rand((n,), 37, 0.1)
    logits = torch.from_numpy(a).double() @ torch.from_numpy(w).double().t() + torch.from_numpy(bias).double()
    top2 = logits.topk(2, -1).values
    clear = ((top2[:, 0] - top2[:, 1]) > 1e-4).numpy()
    ids = raw.vocab_argmax(a, w, bias, precision="bf16x3")
    assert np.array_equal(ids[clear], logits.argmax(-1).numpy()[clear])


def test_bf16x3_is_much_closer_than_bf16(raw):
    a, w, bias = _rand((256, 512), 5), _rand((512, 512), 6, 512 ** -0.5), _rand((512,), 7)
    ref = torch.from_numpy(a).double() @ torch.from_numpy(w).double().t() + torch.from_numpy(bias).double()
    e3, e1 = rel_err(raw.linear(a, w, bias, precision="bf16x3"), ref), rel_err(raw.linear(a, w, bias, precision="bf16"), ref)
    assert e3 * 50 < e1


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("m,n", [(200, 5037), (131, 60515)])
def test_vocab_argmax_never_materialises_logits(raw, gemm_kernel, vocab_mode, precision, m, n):
    if precision == "fp32" and (gemm_kernel == "2cta" or vocab_mode == "full"):
        pytest.skip("fp32 mode does not use the tensor-core kernels")
    a, w, bias = _rand((m, 512), 8), _rand((n, 512), 9, 512 ** -0.5), _rand((n,), 10, 0.1)
    logits = torch.from_numpy(a).double() @ torch.from_numpy(w).double().t() + torch.from_numpy(bias).double()
    top2 = logits.topk(2, -1).values
    clear = ((top2[:, 0] - top2[:, 1]) > 1e-4).numpy()
    ids = raw.vocab_argmax(a, w, bias, precision=precision)
    assert ids.min() >= 0 and ids.max() < n
    assert np.array_equal(ids[clear], logits.argmax(-1).numpy()[clear])


def _argmax_f64(a, w, bias, chunk=500):
    """fp64 argmax and top-2 margin, a block of rows at a time (the full logit matrix would be gigabytes)."""
    wd, bd = torch.from_numpy(w).double().t().contiguous(), torch.from_numpy(bias).double()
    ids, margin, top = [], [], []
    for r in range(0, a.shape[0], chunk):
        lg = torch.from_numpy(a[r:r + chunk]).double() @ wd + bd
        t2 = lg.topk(2, -1).values
        ids.append(lg.argmax(-1)); margin.append(t2[:, 0] - t2[:, 1]); top.append(t2[:, 0].abs())
    return torch.cat(ids).numpy(), torch.cat(margin).numpy(), torch.cat(top).numpy()


@pytest.mark.parametrize("m,n", [(9500, 20011), (19000, 5037)])
def test_vocab_argmax_rescoring_is_fp32_exact(raw, monkeypatch, m, n):
    """Batches of at least half a wave of 256-row blocks take the candidate path (one-product pass on the A-resident
    pair kernel, fp32 rescoring of the surviving columns), so the ids must agree with the fp64 argmax wherever the
    top-2 margin exceeds fp32 rounding — far tighter than the 1e-4 the three-product path is held to — on peaked rows
    (a planted winner) and flat ones (many near-ties inside the one-product error bound)."""
    monkeypatch.delenv("FUNASR_B200_VOCAB", raising=False)
    monkeypatch.delenv("FUNASR_B200_GEMM", raising=False)
    a, w, bias = _rand((m, 512), 41), _rand((n, 512), 42, 512 ** -0.5), _rand((n,), 43, 0.1)
    planted = np.random.default_rng(44).choice(n, size=m // 4, replace=False)
    w[planted] += 0.3 * a[: m // 4] / np.linalg.norm(a[: m // 4], axis=1, keepdims=True)     # clear winners
    want, margin, top = _argmax_f64(a, w, bias)
    clear = margin > 2e-6 * np.maximum(top, 1.0)
    assert clear.mean() > 0.95
    ids = raw.vocab_argmax(a, w, bias, precision="bf16x3")
    assert np.array_equal(ids[clear], want[clear])
    monkeypatch.setenv("FUNASR_B200_GEMM_AR", "0")               # same lists from the general pair kernel's epilogue
    ids = raw.vocab_argmax(a, w, bias, precision="bf16x3")
    assert np.array_equal(ids[clear], want[clear])


@pytest.mark.parametrize("n", [700, 1100])
def test_vocab_argmax_overflowing_lists_take_the_second_chance_pass(raw, monkeypatch, n):
    """Identical weight rows: every column is a candidate, every list overflows, and the gated three-product pass
    has to produce the ids (lowest index on ties, like torch.argmax); n = 1100 goes through the A-resident kernel."""
    monkeypatch.delenv("FUNASR_B200_VOCAB", raising=False)
    monkeypatch.delenv("FUNASR_B200_GEMM", raising=False)
    m = 9500
    a = _rand((m, 512), 45)
    w = np.tile(_rand((1, 512), 46, 512 ** -0.5), (n, 1))
    bias = np.zeros((n,), np.float32)
    assert (raw.vocab_argmax(a, w, bias, precision="bf16x3") == 0).all()
    bias[300] = 1.0
    bias[650] = 1.0
    assert (raw.vocab_argmax(a, w, bias, precision="bf16x3") == 300).all()
    # half of the rows peaked (short lists), half flat (overflow): both kinds in one call
    w2 = w.copy()
    w2[17] += 0.5 * a[0] / np.linalg.norm(a[0])
    a2 = a.copy()
    a2[: m // 2] = a[0]
    bias[:] = 0.0
    ids = raw.vocab_argmax(a2, w2, bias, precision="bf16x3")
    assert (ids[: m // 2] == 17).all()


@pytest.mark.parametrize("m", [300, 9500])
def test_vocab_argmax_all_logits_negative_last_group_empty(raw, vocab_mode, m):
    """N = 300: the second 128-column group of the last tile lies wholly past N.  Its partial-maximum slot must not
    hold stale memory that beats a row whose logits are all negative."""
    n = 300
    a, w = _rand((m, 512), 47), _rand((n, 512), 48, 512 ** -0.5)
    bias = np.full((n,), -50.0, np.float32)
    want, margin, _ = _argmax_f64(a, w, bias)
    ids = raw.vocab_argmax(a, w, bias, precision="bf16x3")
    clear = margin > 1e-4
    assert np.array_equal(ids[clear], want[clear])


def test_vocab_argmax_ties_pick_first_index(raw, vocab_mode):
    # identical weight rows => exactly tied logits; torch.argmax semantics = lowest index (in the candidate path all
    # 700 columns are candidates: the lists overflow and the rows are rescored over the whole vocabulary)
    a = _rand((64, 512), 11)
    w = np.tile(_rand((1, 512), 12, 512 ** -0.5), (700, 1))
    bias = np.zeros((700,), np.float32)
    for precision in ("fp32", "bf16x3"):
        assert (raw.vocab_argmax(a, w, bias, precision=precision) == 0).all()
    bias[300] = 1.0
    bias[650] = 1.0
    for precision in ("fp32", "bf16x3"):
        assert (raw.vocab_argmax(a, w, bias, precision=precision) == 300).all()


ATT_TOL = {"fp32": 5e-6, "bf16x3": 4e-5}


@pytest.fixture(params=["f32", "planes"])
def att_out(request, monkeypatch):
    """Output form of the tensor-core attention kernel under test: fp32 rows, or the bf16 hi/lo planes every engine
    launch writes for the out-projection (handed back as hi + lo)."""
    monkeypatch.setenv("FUNASR_B200_TEST_ATTN_OUT", request.param)
    return request.param


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("heads,dk,t", [(4, 128, 150), (8, 128, 150), (8, 64, 150), (4, 128, 1001), (8, 64, 333)])
def test_attention_masked_keys(raw, att_out, precision, heads, dk, t):
    b, d = 3, heads * dk
    qkv = _rand((b * t, 3 * d), 13, 0.7)
    kv_len = [t, (t * 2) // 3 + 1, 1]
    got = raw.attention(qkv, b, t, heads, dk, kv_len, precision=precision)
    x = torch.from_numpy(qkv).double().view(b, t, 3, heads, dk)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    s = (q * dk ** -0.5) @ k.transpose(-2, -1)
    mask = (torch.arange(t).view(1, 1, 1, t) < torch.tensor(kv_len).view(b, 1, 1, 1)).double()
    s = s + (mask - 1.0) * 10000.0                      # the reference's additive mask (model_definition.py:72-73)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(b * t, d)
    assert rel_err(got, ref) <= ATT_TOL[precision]
    got_full = raw.attention(qkv, b, t, heads, dk, None, precision=precision)   # unmasked: the CTC head's mask=None
    ref_full = (torch.softmax((q * dk ** -0.5) @ k.transpose(-2, -1), -1) @ v).transpose(1, 2).reshape(b * t, d)
    assert rel_err(got_full, ref_full) <= ATT_TOL[precision]


def test_attention_peaked_scores(raw, att_out):
    """Large |scores| (sharp softmax): the two-pass max must keep exp2 in range on the tensor-core path."""
    b, t, heads, dk = 2, 200, 4, 128
    qkv = _rand((b * t, 3 * heads * dk), 21, 1.0)
    qkv[:, : heads * dk] *= 6.0
    x = torch.from_numpy(qkv).double().view(b, t, 3, heads, dk)
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))
    ref = (torch.softmax((q * dk ** -0.5) @ k.transpose(-2, -1), -1) @ v).transpose(1, 2).reshape(b * t, heads * dk)
    for precision in ("fp32", "bf16x3"):
        assert rel_err(raw.attention(qkv, b, t, heads, dk, None, precision=precision), ref) <= 10 * ATT_TOL[precision]


@pytest.mark.parametrize("heads,dk", [(4, 128), (8, 64)])
def test_attention_outlier_key_beyond_the_first_tile_repeats_the_item_exactly(raw, att_out, heads, dk):
    """The tensor-core kernel takes a row's shift from key tile 0 only and repeats an item with the exact two-pass shift
    when a later key beats that shift by more than 2^60 in weight.  Segment 0: no outlier (fast mode only).  Segment 1:
    key 300 scores ~100 (log2 units) above everything for every query, far past fp32 exp2 range without the exact
    shift.  Segment 2: the same outlier inside tile 0 (fast mode copes: it IS the shift).  Segment 3: outlier at key 700
    but masked out by kv_len."""
    b, t, d = 4, 900, heads * dk
    qkv = _rand((b * t, 3 * d), 51, 0.7)
    x = qkv.reshape(b, t, 3, heads, dk)
    u = _rand((dk,), 52)
    u /= np.linalg.norm(u)
    x[1:, :, 0] += 8.0 * u                                   # every query of segments 1..3 leans along u
    for seg, key in ((1, 300), (2, 17), (3, 700)):
        x[seg, key, 1] = 110.0 * u                           # ... and one key is huge along u
    qkv = np.ascontiguousarray(x.reshape(b * t, 3 * d))
    kv_len = [t, t, t, 650]
    xt = torch.from_numpy(qkv).double().view(b, t, 3, heads, dk)
    q, k, v = (xt[:, :, i].transpose(1, 2) for i in range(3))
    s = (q * dk ** -0.5) @ k.transpose(-2, -1)
    assert float((s[1, :, :, 300] - s[1, :, :, :128].amax(-1)).min()) * 1.4427 > 70.0     # the redo really triggers
    mask = (torch.arange(t).view(1, 1, 1, t) < torch.tensor(kv_len).view(b, 1, 1, 1)).double()
    ref = (torch.softmax(s + (mask - 1.0) * 10000.0, -1) @ v).transpose(1, 2).reshape(b * t, d)
    for precision in ("fp32", "bf16x3"):
        got = raw.attention(qkv, b, t, heads, dk, kv_len, precision=precision)
        assert np.isfinite(got).all()
        assert rel_err(got, ref) <= 10 * ATT_TOL[precision]


@pytest.mark.parametrize("d,eps", [(512, 1e-5), (560, 1e-5), (1024, 1e-12), (512, 1e-12)])
def test_layernorm(raw, d, eps):
    x = _rand((333, d), 14, 3.0) + 0.5
    g, b = _rand((d,), 15, 0.1) + 1.0, _rand((d,), 16, 0.1)
    ref = F.layer_norm(torch.from_numpy(x).double(), (d,), torch.from_numpy(g).double(), torch.from_numpy(b).double(), eps)
    got, planes = raw.layernorm(x, g, b, eps)
    assert rel_err(got, ref) <= 2e-6
    assert np.abs(planes - got).max() <= 2.0 ** -15 * np.abs(got).max()


def test_fsmn_memory_block(raw):
    b, t = 3, 77
    v, w = _rand((b, t, 512), 17), _rand((512, 11), 18, 0.3)
    resid = _rand((b, t, 512), 19)
    tv = [77, 40, 3]
    m = (torch.arange(t).view(1, t, 1) < torch.tensor(tv).view(b, 1, 1)).double()
    vm = torch.from_numpy(v).double() * m
    conv = F.conv1d(F.pad(vm.transpose(1, 2), (5, 5)), torch.from_numpy(w).double().unsqueeze(1), groups=512).transpose(1, 2)
    ref = conv + vm
    assert rel_err(raw.fsmn(v, w, tv), ref) <= 2e-6
    assert rel_err(raw.fsmn(v, w, tv, resid), ref + torch.from_numpy(resid).double()) <= 2e-6


def test_fsmn_streaming_kernel_matches_the_strip_kernel_bit_for_bit(raw, monkeypatch):
    """The persistent double-buffered kernel (bulk copies into shared memory, FUNASR_B200_FSMN=stream) against the
    register-strip kernel, on enough strips that every CTA walks several of them: full, ragged, one-frame and empty-tail segments."""
    b, t = 40, 333
    v, w = _rand((b, t, 512), 27), _rand((512, 11), 28, 0.3)
    resid = _rand((b, t, 512), 29)
    tv = [t, 1, 5, 16, 17, 100, 332, 11] * 5
    got = raw.fsmn(v, w, tv, resid)
    got0 = raw.fsmn(v, w, tv)
    monkeypatch.setenv("FUNASR_B200_FSMN", "stream")
    assert np.array_equal(got, raw.fsmn(v, w, tv, resid))
    assert np.array_equal(got0, raw.fsmn(v, w, tv))
    m = (torch.arange(t).view(1, t, 1) < torch.tensor(tv).view(b, 1, 1)).double()
    vm = torch.from_numpy(v).double() * m
    conv = F.conv1d(F.pad(vm.transpose(1, 2), (5, 5)), torch.from_numpy(w).double().unsqueeze(1), groups=512).transpose(1, 2)
    assert rel_err(got, conv + vm + torch.from_numpy(resid).double()) <= 2e-6


@pytest.mark.parametrize("precision", ["bf16x3", "bf16", "fp8"])
@pytest.mark.parametrize("m,n,k", [(300, 512, 512), (1001, 512, 2048), (2100, 1024, 256), (77, 512, 128)])
def test_in_place_residual_is_a_tensor_reduction_with_the_same_bits(raw, gemm_kernel, monkeypatch, precision, m, n, k):
    """x = x + A W^T + b with the residual being the output buffer (the out-projection and FFN 2 of every layer): the
    epilogue does not read x, it hands acc + bias to a tensor REDUCTION (cp.reduce.async.bulk.tensor .add) and the memory
    system adds it to x.  One fp32 round-to-nearest add per element, each element exactly once: the result must be
    bit-identical to the read-add-store epilogue that a separate residual buffer takes."""
    if precision == "fp8" and gemm_kernel == "1cta":
        pytest.skip("the fp8 projections always run on the CTA-pair kernel")
    a, w = _rand((m, k), 21), _rand((n, k), 22, k ** -0.5)
    bias, resid = _rand((n,), 23), _rand((m, n), 24)
    separate = raw.linear(a, w, bias, resid=resid, precision=precision)
    monkeypatch.setenv("FUNASR_B200_TEST_INPLACE", "1")
    in_place = raw.linear(a, w, bias, resid=resid, precision=precision)
    assert np.array_equal(separate, in_place)


# ------------------------------------------------------------------------------------ fp8 speed mode

def _e4m3(x: torch.Tensor) -> torch.Tensor:
    """Round to e4m3 (nearest even, saturating at +-448) and back, on the CPU: what the kernels' conversions do."""
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).to(torch.float32)


@pytest.mark.parametrize("m,n,k", [(300, 512, 512), (129, 1536, 560), (515, 2048, 512), (1001, 512, 2048), (260, 256, 128),
                                   (2100, 1024, 256)])
def test_linear_fp8_is_the_quantised_product(raw, m, n, k):
    """FA_PREC_FP8 projections: tcgen05 kind::f8f6f4 on e4m3 activations (converted as they are) and e4m3 weights with a
    per-output-channel scale max|w_row| / 448 (02-Quantize-ONNX.py:41-44, per_channel=True).  The kernel must compute
    exactly that quantised product — the e4m3 x e4m3 products are exact in fp32, so against a float64 product of the
    same quantised operands only the fp32 accumulation order is left — with bias, ReLU, residual and the e4m3 output
    form the next projection reads."""
    a, w = _rand((m, k), 11), _rand((n, k), 12, k ** -0.5)
    bias, resid = _rand((n,), 13), _rand((m, n), 14)
    ta, tw = torch.from_numpy(a), torch.from_numpy(w)
    scale = tw.abs().amax(1) / 448.0
    a8, w8 = _e4m3(ta), _e4m3(tw / scale[:, None])
    ref = (a8.double() @ w8.double().t()) * scale.double() + torch.from_numpy(bias).double()
    got = raw.linear(a, w, bias, precision="fp8")
    assert rel_err(got, ref) <= 2e-6
    # and it is a sane approximation of the fp32 product (3 mantissa bits per operand, K-fold averaging)
    exact = ta.double() @ tw.double().t() + torch.from_numpy(bias).double()
    assert rel_err(got, exact) <= 0.08
    got, out8 = raw.linear(a, w, bias, relu=True, precision="fp8", planes=True)
    assert rel_err(got, torch.relu(ref)) <= 2e-6
    assert np.array_equal(out8, _e4m3(torch.from_numpy(got)).numpy())                  # the e4m3 output is the rounding of the fp32 one
    got = raw.linear(a, w, bias, resid=resid, precision="fp8")
    assert rel_err(got, ref + torch.from_numpy(resid).double()) <= 2e-6
