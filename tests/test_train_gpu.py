"""GPU tier: the training step (src/weinman/train.py) on B200.

Per-op checks compare each C-ABI entry point with a float64 torch (CPU) statement of the same TensorFlow op / gradient;
the end-to-end check compares one whole training step with oracle/train_oracle.py (float64 autograd restatement of
train.py's graph).

Tolerance.  Memory-bound kernels are fp32: asserted to 1e-5 relative.  Contractions are TF32 tensor-core products
(10-bit mantissas, fp32 sums): asserted to 5e-3 of the result's scale per op; gradients of the full step pass through
~20 such contractions forward and backward and are asserted to 3e-2 of each tensor's max |gradient| (measured ~3e-3)."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _lib():
    from cnn_lstm_ctc_ocr_b200 import _lib as L
    return L, L.load(), L.stream_handle()


_KEEP = []   # L.ptr(_t(x)) hands a raw pointer to the library: keep the tensor alive past the call


def _t(a):
    t = torch.tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)
    _KEEP.append(t)
    if len(_KEEP) > 64:
        torch.cuda.synchronize()
        del _KEEP[:32]
    return t


def _close(got, ref, rel, what=""):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    scale = max(np.abs(ref).max(), 1e-30)
    err = np.abs(got - ref).max()
    assert err <= rel * scale, "%s: max err %.3e vs scale %.3e (rel %.2e > %.1e)" % (what, err, scale, err / scale, rel)


def test_transpose_and_planar_pad():
    L, lib, sh = _lib()
    rng = np.random.default_rng(0)
    x = rng.standard_normal((77, 45)).astype(np.float32)
    out = torch.full((45, 80), 7.0, device=DEV)
    L.check(lib.ocr_transpose(L.ptr(_t(x)), 77, 45, 45, L.ptr(out), 80, 0, sh), "t")
    o = out.cpu().numpy()
    assert (o[:, :77] == x.T).all() and (o[:, 77:] == 7.0).all()
    out = torch.full((45, 80), 7.0, device=DEV)      # shifted: out[c, r] = x[r - 5, c]
    L.check(lib.ocr_transpose(L.ptr(_t(x)), 77, 45, 45, L.ptr(out), 80, -5, sh), "t")
    o = out.cpu().numpy()
    assert (o[:, 5:77] == x.T[:, :72]).all() and (o[:, :5] == 0).all()
    B, H, W, C = 2, 3, 5, 36
    a = rng.standard_normal((B, H, W, C)).astype(np.float32)
    Wp = lib.ocr_planar_pad_pitch(W)
    assert Wp == 8
    R = B * (H + 2) * Wp
    out = torch.full((3, C, R), 7.0, device=DEV)
    L.check(lib.ocr_nhwc_to_planar_pad(L.ptr(_t(a)), B, H, W, C, L.ptr(out), R, 3, C * R, sh), "p")
    ref = np.zeros((B, H + 2, Wp, C), np.float32)
    ref[:, 1:H + 1, 1:W + 1] = a
    flat = ref.reshape(R, C).T
    o = out.cpu().numpy()
    assert (o[1] == flat).all()
    assert (o[0][:, 1:] == flat[:, :-1]).all() and (o[0][:, 0] == 0).all()      # copy 0 holds pixel r-1 at r
    assert (o[2][:, :-1] == flat[:, 1:]).all() and (o[2][:, -1] == 0).all()     # copy 2 holds pixel r+1 at r


@pytest.mark.parametrize("stack", [1, 0])
@pytest.mark.parametrize("M,N,R,shifts", [(40, 24, 1000, [0]), (256, 128, 33000, [0]), (70, 300, 5000, [-4, 0, 12]),
                                          (32, 32, 70001, [-68, -68, -68, 0, 0, 0, 68, 68, 68]),
                                          (64, 64, 20000, [-68, -64, -60, -4, 0, 4, 60, 64, 68]), (48, 40, 9000, [-4, 0, 4, 8, 12]),
                                          (40, 24, 1000, [0, 4, 8])])
def test_wgrad_gemm(M, N, R, shifts, stack):
    """stack = 1 (default): views of at most 64 rows are stacked in one 128-row operand tile (one MMA contracts several taps);
    stack = 0: one tile per view."""
    L, lib, sh = _lib()
    L.check(lib.ocr_debug_gemm_tma_store(1 if stack else 3), "stack")
    rng = np.random.default_rng(1)
    ld = (R + 3) // 4 * 4
    At = np.zeros((M, ld), np.float32); At[:, :R] = rng.standard_normal((M, R))
    Wt = np.zeros((N, ld), np.float32); Wt[:, :R] = rng.standard_normal((N, R))
    At[:, R:] = 1e6                                          # the pad is outside the tensor map: must never be read
    nb = len(shifts)
    need = ctypes.c_size_t(0)
    L.check(lib.ocr_gemm_wgrad_scratch_bytes(M, N, R, nb, ctypes.byref(need)), "s")
    scr = torch.empty(need.value, dtype=torch.uint8, device=DEV)
    D = torch.zeros((nb, M, N), device=DEV)
    arr = (ctypes.c_int32 * nb)(*shifts)
    L.check(lib.ocr_gemm_tf32_wgrad(L.ptr(_t(At)), ld, L.ptr(_t(Wt)), ld, L.ptr(D), N, M * N, M, N, R, nb, arr, None, 0, L.ptr(scr), need.value, sh), "w")
    bad = (ctypes.c_int32 * nb)(*([1] * nb))      # box origins that are not 16-byte aligned are refused, not faulted on
    assert lib.ocr_gemm_tf32_wgrad(L.ptr(_t(At)), ld, L.ptr(_t(Wt)), ld, L.ptr(D), N, M * N, M, N, R, nb, bad, None, 0, L.ptr(scr), need.value, sh) != 0
    A64, W64 = At[:, :R].astype(np.float64), Wt[:, :R].astype(np.float64)
    for i, s in enumerate(shifts):
        Ash = np.zeros_like(A64)
        if s >= 0:
            Ash[:, :R - s] = A64[:, s:]
        else:
            Ash[:, -s:] = A64[:, :R + s]
        ref = Ash @ W64.T
        _close(D[i].cpu().numpy(), ref, 3e-3, "wgrad shift %d" % s)
    L.check(lib.ocr_debug_gemm_tma_store(1), "stack")


def test_batch_norm_train_and_backward():
    L, lib, sh = _lib()
    rng = np.random.default_rng(2)
    rows, C = 5000, 64
    y = (rng.standard_normal((rows, C)) * rng.uniform(0.5, 3, C) + rng.normal(0, 2, C)).astype(np.float32)
    gamma, beta = rng.uniform(0.5, 1.5, C).astype(np.float32), rng.normal(0, 0.3, C).astype(np.float32)
    mm, mv = rng.normal(0, 1, C).astype(np.float32), rng.uniform(0.5, 2, C).astype(np.float32)
    g = rng.standard_normal((rows, C)).astype(np.float32)
    yt = torch.tensor(y, dtype=torch.float64, requires_grad=True)
    gt, bt = torch.tensor(gamma, dtype=torch.float64, requires_grad=True), torch.tensor(beta, dtype=torch.float64, requires_grad=True)
    mean, var = yt.mean(0), yt.var(0, unbiased=False)
    out_ref = torch.relu(gt * (yt - mean) / torch.sqrt(var + 1e-3) + bt)
    out_ref.backward(torch.tensor(g, dtype=torch.float64))
    sums = torch.empty(2 * C, dtype=torch.float64, device=DEV)
    d_y, d_mean, d_is, d_out = _t(y), torch.empty(C, device=DEV), torch.empty(C, device=DEV), torch.empty((rows, C), device=DEV)
    d_mm, d_mv, d_g, d_b = _t(mm), _t(mv), _t(gamma), _t(beta)
    L.check(lib.ocr_bn_batch_sums(L.ptr(d_y), rows, C, L.ptr(sums), sh), "sums")
    L.check(lib.ocr_bn_finalize(L.ptr(sums), rows, C, 1e-3, 0.99, L.ptr(d_mean), L.ptr(d_is), L.ptr(d_mm), L.ptr(d_mv), sh), "fin")
    L.check(lib.ocr_bn_relu_apply(L.ptr(d_y), rows, C, L.ptr(d_mean), L.ptr(d_is), L.ptr(d_g), L.ptr(d_b), L.ptr(d_out), sh), "apply")
    _close(d_mean.cpu().numpy(), mean.detach().numpy(), 1e-6, "mean")
    _close(d_is.cpu().numpy(), (1 / torch.sqrt(var + 1e-3)).detach().numpy(), 1e-6, "inv_std")
    _close(d_out.cpu().numpy(), out_ref.detach().numpy(), 1e-5, "bn out")
    _close(d_mm.cpu().numpy(), 0.99 * mm + 0.01 * mean.detach().numpy(), 1e-6, "moving mean")
    _close(d_mv.cpu().numpy(), 0.99 * mv + 0.01 * var.detach().numpy() * rows / (rows - 1), 1e-6, "moving var")
    dgam, dbet, dy, dg = torch.empty(C, device=DEV), torch.empty(C, device=DEV), torch.empty((rows, C), device=DEV), _t(g)
    L.check(lib.ocr_bn_relu_bwd_sums(L.ptr(d_y), L.ptr(dg), rows, C, L.ptr(d_mean), L.ptr(d_is), L.ptr(d_g), L.ptr(d_b), L.ptr(sums), L.ptr(dgam), L.ptr(dbet), sh), "bs")
    L.check(lib.ocr_bn_relu_bwd_apply(L.ptr(d_y), L.ptr(dg), rows, rows, C, L.ptr(d_mean), L.ptr(d_is), L.ptr(d_g), L.ptr(d_b), L.ptr(sums), L.ptr(dy), sh), "ba")
    _close(dgam.cpu().numpy(), gt.grad.numpy(), 1e-5, "dgamma")
    _close(dbet.cpu().numpy(), bt.grad.numpy(), 1e-5, "dbeta")
    _close(dy.cpu().numpy(), yt.grad.numpy(), 2e-5, "dy")
    # the same pass also yields the bias gradient of the convolution in front (column sums of dy): same dy bits, sums as ocr_colsum's
    dy2, db2, db_ref = torch.empty((rows, C), device=DEV), torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    scr = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    L.check(lib.ocr_bn_relu_bwd_apply_bias(L.ptr(d_y), L.ptr(dg), rows, rows, C, L.ptr(d_mean), L.ptr(d_is), L.ptr(d_g), L.ptr(d_b), L.ptr(sums), L.ptr(dy2),
                                           L.ptr(db2), L.ptr(scr), sh), "bab")
    L.check(lib.ocr_colsum(L.ptr(dy), rows, C, C, L.ptr(db_ref), L.ptr(scr), sh), "colsum")
    assert torch.equal(dy2, dy)
    want = dy.double().sum(0).cpu().numpy()
    scale = float(dy.double().abs().sum(0).max())       # the sums cancel almost completely (sum of xhat = 0): compare on the scale of the terms
    assert np.abs(db2.cpu().numpy() - want).max() <= 1e-6 * scale and np.abs(db_ref.cpu().numpy() - want).max() <= 1e-6 * scale


@pytest.mark.parametrize("B,H,W,C,sw", [(2, 30, 37, 32, 2), (3, 15, 21, 64, 1), (2, 7, 62, 128, 1), (1, 2, 2, 4, 2), (2, 5, 9, 8, 2)])
def test_bn_relu_apply_pool_equals_apply_then_pool(B, H, W, C, sw):
    """ocr_bn_relu_apply_pool == ocr_bn_relu_apply followed by ocr_maxpool(2, 2, 2, sw), bit for bit, and nothing else is written."""
    L, lib, sh = _lib()
    rng = np.random.default_rng(B * 100 + W)
    y = _t(rng.standard_normal((B, H, W, C)) * 2)
    mean, inv_std = _t(rng.normal(0, 0.5, C)), _t(rng.uniform(0.5, 2, C))
    gamma, beta = _t(rng.uniform(0.5, 1.5, C)), _t(rng.normal(0, 0.3, C))
    Hp, Wp = (H - 2) // 2 + 1, (W - 2) // sw + 1
    z_ref, p_ref = torch.empty((B, H, W, C), device=DEV), torch.empty((B, Hp, Wp, C), device=DEV)
    L.check(lib.ocr_bn_relu_apply(L.ptr(y), B * H * W, C, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta), L.ptr(z_ref), sh), "apply")
    L.check(lib.ocr_maxpool(L.ptr(z_ref), B, H, W, C, 2, 2, 2, sw, L.ptr(p_ref), sh), "pool")
    z = torch.full((B * H * W * C + 16,), float("nan"), device=DEV)
    pz = torch.full((B * Hp * Wp * C + 16,), float("nan"), device=DEV)
    L.check(lib.ocr_bn_relu_apply_pool(L.ptr(y), B, H, W, C, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta), L.ptr(z), sw, L.ptr(pz), sh), "fused")
    torch.cuda.synchronize()
    assert torch.equal(z[:-16].view(B, H, W, C), z_ref) and torch.isnan(z[-16:]).all()
    assert torch.equal(pz[:-16].view(B, Hp, Wp, C), p_ref) and torch.isnan(pz[-16:]).all()


@pytest.mark.parametrize("B,H,W,C,sw", [(2, 30, 37, 32, 2), (3, 15, 21, 64, 1), (2, 7, 62, 128, 1), (1, 2, 2, 4, 2), (2, 5, 9, 8, 2), (1, 3, 126, 256, 1)])
def test_pool_arg_path_equals_separate_pool_gradient(B, H, W, C, sw):
    """Pooled batch-norm layer without the full-size activation: ocr_bn_relu_apply_pool_arg gives the bits of apply + maxpool, and
    the two backward passes that form MaxPoolGrad from the stored arguments of the maxima give the bits (dy) / the sums (float64
    accumulators, 1e-6) of ocr_maxpool_bwd followed by ocr_bn_relu_bwd_sums / _apply_bias.  Values are quantised so that windows
    hold ties (TensorFlow's first-maximum rule decides) and whole windows of zeros."""
    L, lib, sh = _lib()
    rng = np.random.default_rng(B * 1000 + W * 10 + C)
    y = _t(np.round(rng.standard_normal((B, H, W, C)) * 2) / 2)
    mean, inv_std = _t(np.round(rng.normal(0, 0.5, C) * 2) / 2), _t(rng.choice([0.5, 1.0, 2.0], C))
    gamma, beta = _t(rng.choice([0.5, 1.0, 1.5], C)), _t(np.round(rng.normal(0, 0.3, C) * 2) / 2)
    Hp, Wp = (H - 2) // 2 + 1, (W - 2) // sw + 1
    rows = B * H * W
    z_ref, p_ref = torch.empty((B, H, W, C), device=DEV), torch.empty((B, Hp, Wp, C), device=DEV)
    L.check(lib.ocr_bn_relu_apply(L.ptr(y), rows, C, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta), L.ptr(z_ref), sh), "apply")
    L.check(lib.ocr_maxpool(L.ptr(z_ref), B, H, W, C, 2, 2, 2, sw, L.ptr(p_ref), sh), "pool")
    pz = torch.full((B * Hp * Wp * C + 16,), float("nan"), device=DEV)
    arg = torch.full((B * Hp * Wp * (C // 4) + 16,), 0xEE, dtype=torch.uint8, device=DEV)
    L.check(lib.ocr_bn_relu_apply_pool_arg(L.ptr(y), B, H, W, C, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta), sw, L.ptr(pz), L.ptr(arg), sh), "fwd")
    torch.cuda.synchronize()
    assert torch.equal(pz[:-16].view(B, Hp, Wp, C), p_ref) and torch.isnan(pz[-16:]).all() and (arg[-16:] == 0xEE).all()
    assert (z_ref == 0).float().mean() > 0.2      # the case does hold zero windows / ties
    # backward, reference: separate pool gradient, then the two batch-norm passes
    dpool = _t(rng.standard_normal((B, Hp, Wp, C)))
    da = torch.empty_like(z_ref)
    L.check(lib.ocr_maxpool_bwd(L.ptr(z_ref), L.ptr(dpool), B, H, W, C, 2, 2, 2, sw, L.ptr(da), sh), "pool bwd")
    sums_ref = torch.empty(2 * C, dtype=torch.float64, device=DEV)
    dg_ref, db_ref = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    L.check(lib.ocr_bn_relu_bwd_sums(L.ptr(y), L.ptr(da), rows, C, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta), L.ptr(sums_ref),
                                     L.ptr(dg_ref), L.ptr(db_ref), sh), "sums")
    dy_ref, dbias_ref = torch.empty_like(z_ref), torch.empty(C, device=DEV)
    scratch = torch.zeros(2 * C, dtype=torch.float64, device=DEV)
    L.check(lib.ocr_bn_relu_bwd_apply_bias(L.ptr(y), L.ptr(da), rows, rows, C, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta), L.ptr(sums_ref),
                                           L.ptr(dy_ref), L.ptr(dbias_ref), L.ptr(scratch), sh), "apply bias")
    # fused
    sums = torch.empty(2 * C, dtype=torch.float64, device=DEV)
    dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
    L.check(lib.ocr_bn_relu_bwd_sums_pool(L.ptr(y), L.ptr(dpool), L.ptr(arg), B, H, W, C, sw, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma), L.ptr(beta),
                                          L.ptr(sums), L.ptr(dg), L.ptr(db), sh), "sums pool")
    dy = torch.full((rows * C + 16,), float("nan"), device=DEV)
    dbias = torch.empty(C, device=DEV)
    L.check(lib.ocr_bn_relu_bwd_apply_bias_pool(L.ptr(y), L.ptr(dpool), L.ptr(arg), B, H, W, C, sw, rows, L.ptr(mean), L.ptr(inv_std), L.ptr(gamma),
                                                L.ptr(beta), L.ptr(sums_ref), L.ptr(dy), L.ptr(dbias), L.ptr(scratch), sh), "apply bias pool")
    torch.cuda.synchronize()
    scale = float(sums_ref.abs().max()) + 1.0
    assert float((sums - sums_ref).abs().max()) <= 1e-6 * scale
    assert float((dg - dg_ref).abs().max()) <= 1e-6 * scale and float((db - db_ref).abs().max()) <= 1e-6 * scale
    assert torch.equal(dy[:-16].view(B, H, W, C), dy_ref) and torch.isnan(dy[-16:]).all()
    assert float((dbias - dbias_ref).abs().max()) <= 1e-5 * (float(dbias_ref.abs().max()) + 1.0)


def test_relu_bias_colsum_pool_gradients():
    L, lib, sh = _lib()
    rng = np.random.default_rng(3)
    scr = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    rows, C = 3001, 96
    out = np.maximum(rng.standard_normal((rows, C)), 0).astype(np.float32)
    g = rng.standard_normal((rows, C)).astype(np.float32)
    dy, db = torch.empty((rows, C), device=DEV), torch.empty(C, device=DEV)
    L.check(lib.ocr_relu_bwd_bias(L.ptr(_t(out)), L.ptr(_t(g)), rows, C, L.ptr(dy), L.ptr(db), L.ptr(scr), sh), "rb")
    ref = g * (out > 0)
    assert (dy.cpu().numpy() == ref).all()
    _close(db.cpu().numpy(), ref.astype(np.float64).sum(0), 1e-5, "dbias")
    cs = torch.empty(C, device=DEV)
    L.check(lib.ocr_colsum(L.ptr(_t(g)), rows, C, C, L.ptr(cs), L.ptr(scr), sh), "cs")
    _close(cs.cpu().numpy(), g.astype(np.float64).sum(0), 1e-5, "colsum")
    dz = torch.empty((rows, C), device=DEV)
    L.check(lib.ocr_relu_bwd(L.ptr(_t(out)), L.ptr(_t(g)), rows * C, L.ptr(dz), sh), "r")
    assert (dz.cpu().numpy() == ref).all()
    # max-pool gradients; post-ReLU inputs (exact ties at zero) like the network's
    for (B, H, W, Cc, ph, pw, s_h, s_w) in [(2, 30, 37, 32, 2, 2, 2, 2), (2, 15, 21, 64, 2, 2, 2, 1), (3, 7, 9, 8, 2, 2, 2, 1)]:
        a = np.maximum(rng.standard_normal((B, H, W, Cc)), 0).astype(np.float32)
        at = torch.tensor(a, dtype=torch.float64).permute(0, 3, 1, 2).requires_grad_(True)
        p = F.max_pool2d(at, (ph, pw), (s_h, s_w))
        gp = rng.standard_normal(tuple(p.shape)).astype(np.float32)
        p.backward(torch.tensor(gp, dtype=torch.float64))
        din = torch.empty((B, H, W, Cc), device=DEV)
        L.check(lib.ocr_maxpool_bwd(L.ptr(_t(a)), L.ptr(_t(np.transpose(gp, (0, 2, 3, 1)))), B, H, W, Cc, ph, pw, s_h, s_w, L.ptr(din), sh), "mp")
        ref = at.grad.permute(0, 2, 3, 1).numpy()
        # where the input is positive the routing is unambiguous; at exact-zero ties both sides pick the first maximum
        _close(din.cpu().numpy(), ref, 1e-6, "maxpool_bwd")
    B, H, W, Cc = 2, 3, 11, 16
    a = np.maximum(rng.standard_normal((B, H, W, Cc)), 0).astype(np.float32)
    at = torch.tensor(a, dtype=torch.float64, requires_grad=True)
    seq = at.max(dim=1).values.permute(1, 0, 2)       # [W,B,C]
    gs = rng.standard_normal((W, B, Cc)).astype(np.float32)
    din = torch.empty((B, H, W, Cc), device=DEV)
    L.check(lib.ocr_rows_max_to_seq_bwd(L.ptr(_t(a)), L.ptr(_t(gs)), B, H, W, Cc, L.ptr(din), sh), "rm")
    got = din.cpu().numpy()
    pos = a.max(axis=1, keepdims=True) > 0            # unique maxima (ties only occur at zero)
    seq.backward(torch.tensor(gs, dtype=torch.float64))
    assert np.allclose((got * pos), (at.grad.numpy() * pos), atol=1e-6)
    assert np.allclose(got.sum(axis=1), np.transpose(gs, (1, 0, 2)), atol=1e-6)


def test_conv_gradients():
    """d input (rotated-filter convolution), d kernel (nine shifted contractions), conv1's d kernel."""
    L, lib, sh = _lib()
    rng = np.random.default_rng(4)
    B, H, W, C, Co = 2, 7, 19, 32, 64
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    w = (rng.standard_normal((3, 3, C, Co)) * 0.1).astype(np.float32)
    dy = rng.standard_normal((B, H, W, Co)).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64).permute(0, 3, 1, 2).requires_grad_(True)
    wt = torch.tensor(w, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xt, wt.permute(3, 2, 0, 1), padding=1)
    y.backward(torch.tensor(dy, dtype=torch.float64).permute(0, 3, 1, 2))
    wf, wd = torch.empty((Co, 9 * C), device=DEV), torch.empty((C, 9 * Co), device=DEV)
    L.check(lib.ocr_conv_filter_layouts(L.ptr(_t(w)), C, Co, L.ptr(wf), L.ptr(wd), sh), "fl")
    assert (wf.cpu().numpy() == np.transpose(w.reshape(9 * C, Co))).all()
    zero = torch.zeros(256, device=DEV)
    dx = torch.empty((B, H, W, C), device=DEV)
    L.check(lib.ocr_conv3x3_same(L.ptr(_t(dy)), B, H, W, Co, L.ptr(wd), L.ptr(zero), C, 0, L.ptr(dx), sh), "dgrad")
    _close(dx.cpu().numpy(), xt.grad.permute(0, 2, 3, 1).numpy(), 3e-3, "conv dgrad")
    # weight gradient through the Trainer helper (planar pad + 9 shifts)
    from cnn_lstm_ctc_ocr_b200 import train
    tr = object.__new__(train.Trainer)
    tr.lib, tr.device, tr.wscratch = lib, torch.device(DEV), None
    for blocked in (True, False):       # K-blocked planar operands (default) and the plain planar copies
        tr.blocked_planar = blocked
        tr.grads = {"convnet/convX/kernel": torch.zeros((3, 3, C, Co), device=DEV)}
        tr._conv_wgrad(_t(x), _t(dy), "convX")
        _close(tr.grads["convnet/convX/kernel"].cpu().numpy(), wt.grad.numpy(), 3e-3, "conv wgrad (blocked=%s)" % blocked)
    # conv1: one input channel, 'valid'
    img = rng.integers(0, 256, (B, 12, 23)).astype(np.uint8)
    Co1 = 32
    dy1 = rng.standard_normal((B, 10, 21, Co1)).astype(np.float32)
    it = (torch.tensor(img, dtype=torch.float64) / 255.0 - 0.5)[:, None]
    w1 = torch.zeros((3, 3, 1, Co1), dtype=torch.float64, requires_grad=True)
    F.conv2d(it, w1.permute(3, 2, 0, 1)).backward(torch.tensor(dy1, dtype=torch.float64).permute(0, 3, 1, 2))
    dw1 = torch.empty((3, 3, 1, Co1), device=DEV)
    dimg = torch.tensor(img, device=DEV)
    scr = torch.zeros(1 << 16, dtype=torch.uint8, device=DEV)
    L.check(lib.ocr_conv1_wgrad(L.ptr(dimg), 1, B, 12, 23, L.ptr(_t(dy1)), Co1, L.ptr(dw1), L.ptr(scr), sh), "c1")
    _close(dw1.cpu().numpy(), w1.grad.numpy(), 1e-5, "conv1 wgrad")


@pytest.mark.parametrize("B,H,W,C,Co", [(3, 5, 31, 32, 32), (2, 6, 63, 64, 128), (2, 3, 126, 128, 128), (1, 30, 254, 32, 32)])
def test_conv_wgrad_blocked_operands(B, H, W, C, Co):
    """Filter gradient over K-blocked planar operands (pitch = W + 1 rounded to 32: the zero column between rows is shared;
    W = 31 and 63 give pitch == W + 1 exactly) against float64 autograd, and against the plain planar path."""
    from cnn_lstm_ctc_ocr_b200 import train
    L, lib, sh = _lib()
    rng = np.random.default_rng(B * 100 + W)
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    dy = rng.standard_normal((B, H, W, Co)).astype(np.float32)
    xt = torch.tensor(x, dtype=torch.float64).permute(0, 3, 1, 2)
    wt = torch.zeros((3, 3, C, Co), dtype=torch.float64, requires_grad=True)
    F.conv2d(xt, wt.permute(3, 2, 0, 1), padding=1).backward(torch.tensor(dy, dtype=torch.float64).permute(0, 3, 1, 2))
    tr = object.__new__(train.Trainer)
    tr.lib, tr.device, tr.wscratch = lib, torch.device(DEV), None
    got = {}
    for blocked in (True, False):
        tr.blocked_planar = blocked
        tr.grads = {"convnet/convX/kernel": torch.zeros((3, 3, C, Co), device=DEV)}
        tr._conv_wgrad(_t(x), _t(dy), "convX")
        got[blocked] = tr.grads["convnet/convX/kernel"].cpu().numpy()
        _close(got[blocked], wt.grad.numpy(), 3e-3, "conv wgrad (blocked=%s)" % blocked)
    _close(got[True], got[False], 1e-4, "blocked vs planar")      # same TF32 products, different split-K partition


def test_adam_matches_oracle():
    from oracle import train_oracle as to
    L, lib, sh = _lib()
    rng = np.random.default_rng(5)
    n = 1003
    p, g = rng.standard_normal(n).astype(np.float32), (rng.standard_normal(n) * 1e-3).astype(np.float32)
    m, v = (rng.standard_normal(n) * 1e-3).astype(np.float32), (rng.uniform(0, 1e-6, n)).astype(np.float32)
    t, lr = 7, to.learning_rate(6)
    pr, mr, vr = to.adam_step(p.astype(np.float64), g.astype(np.float64) * 0.5, m.astype(np.float64), v.astype(np.float64), t, lr)
    dp, dg, dm, dv = _t(p), _t(g), _t(m), _t(v)
    lr_t = lr * np.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
    L.check(lib.ocr_adam_step(L.ptr(dp), L.ptr(dg), L.ptr(dm), L.ptr(dv), n, lr_t, None, 0.9, 0.999, 1e-8, 0.5, sh), "adam")
    assert np.abs(dp.cpu().numpy() - pr).max() < 2.5e-7     # one float32 ulp of |p| ~ 1-2
    _close(dm.cpu().numpy(), mr, 1e-6, "m")
    _close(dv.cpu().numpy(), vr, 1e-6, "v")


def _small_problem(seed=0, B=4, W=44, sizes=(32, 32), classes=19, cell="lstm"):
    """B = 4: frame shifts of the transposed outputs are aligned TMA offsets; B = 3 takes the shifted-copy path."""
    from oracle import model_oracle as mo
    rng = np.random.default_rng(seed)
    params = mo.init_params(seed, cell, sizes, classes, np.float64, randomize_bn=True)
    for k in list(params):       # make the recurrence and the biases matter
        if "_cell/" in k and k.endswith("kernel"):
            params[k] = params[k] * 8
        if k.endswith("bias"):
            params[k] = params[k] + rng.normal(0, 0.05, params[k].shape)
    img = rng.integers(0, 256, (B, 32, W, 1)).astype(np.uint8)
    widths = np.array([W, W - 7, W - 11, W - 2][:B])
    labels = [[1, 2, 3], [4], [5, 5, 0], [7, 8]][:B]
    return params, img, widths, labels


@pytest.mark.parametrize("H,path", [(16, 0), (32, 0), (32, 1), (32, 2), (64, 3)])
def test_lstm_layer_train_forward_backward(H, path):
    """ocr_birnn_lstm_train_fwd / _bwd against autograd through the oracle's bidirectional_dynamic_rnn restatement.
    H = 32, path 0: forward and BPTT frames run in the persistent cooperative kernels; path 1 / H = 16: one launch pair per
    frame; path 2: persistent forward only; path 3: persistent BPTT forced."""
    from oracle import train_oracle as to
    L, lib, sh = _lib()
    rng = np.random.default_rng(6)
    T, B, I = 9, 5, 24
    sl = np.array([9, 4, 7, 1, 9], np.int32)
    L.check(lib.ocr_birnn_set_path(path), "path")
    x = rng.standard_normal((T, B, I))
    ks = {d: rng.standard_normal((I + H, 4 * H)) * 0.3 for d in ("fw", "bw")}
    bs = {d: rng.standard_normal(4 * H) * 0.1 for d in ("fw", "bw")}
    gout = rng.standard_normal((T, B, 2 * H))
    tp = {}
    for d in ("fw", "bw"):
        tp["p/%s/lstm_cell/kernel" % d] = torch.tensor(ks[d], requires_grad=True)
        tp["p/%s/lstm_cell/bias" % d] = torch.tensor(bs[d], requires_grad=True)
    xt = torch.tensor(x, requires_grad=True)
    ref = torch.cat([to._run_direction(xt, sl, tp, "p/%s/" % d, "lstm", H, rev) for d, rev in (("fw", False), ("bw", True))], dim=2)
    ref.backward(torch.tensor(gout))
    wx = _t(np.concatenate([ks["fw"][:I].T, ks["bw"][:I].T], 0))
    wh = _t(np.concatenate([ks["fw"][I:].T, ks["bw"][I:].T], 0))
    bias = _t(np.concatenate([bs["fw"], bs["bw"]]))
    wh_rows = _t(np.concatenate([ks["fw"][I:], ks["bw"][I:]], 0))
    need = ctypes.c_size_t(0)
    L.check(lib.ocr_birnn_lstm_train_workspace_bytes(T, B, H, ctypes.byref(need)), "ws")
    ws = torch.empty(need.value, dtype=torch.uint8, device=DEV)
    out, gates, cs = torch.empty((T, B, 2 * H), device=DEV), torch.empty((T * B, 8 * H), device=DEV), torch.empty((T, B, 2 * H), device=DEV)
    dsl = torch.tensor(sl, device=DEV)
    L.check(lib.ocr_birnn_lstm_train_fwd(L.ptr(_t(x)), T, B, I, H, L.ptr(dsl), L.ptr(wx), L.ptr(wh), L.ptr(bias), L.ptr(out), L.ptr(gates), L.ptr(cs),
                                         L.ptr(ws), need.value, sh), "fwd")
    _close(out.cpu().numpy(), ref.detach().numpy(), 5e-3, "lstm out")
    L.check(lib.ocr_birnn_lstm_bwd(L.ptr(_t(gout)), T, B, H, L.ptr(dsl), L.ptr(gates), L.ptr(cs), L.ptr(wh_rows), L.ptr(ws), need.value, sh), "bwd")
    dG = gates.cpu().numpy().astype(np.float64)           # [T*B, 8H]
    X = x.reshape(T * B, I)
    for d, dn in enumerate(("fw", "bw")):
        dGd = dG[:, d * 4 * H:(d + 1) * 4 * H]
        _close(dGd.sum(0), tp["p/%s/lstm_cell/bias" % dn].grad.numpy(), 1e-2, "dbias " + dn)
        _close(X.T @ dGd, tp["p/%s/lstm_cell/kernel" % dn].grad.numpy()[:I], 1e-2, "dWx " + dn)
    dx = dG[:, :4 * H] @ ks["fw"][:I].T + dG[:, 4 * H:] @ ks["bw"][:I].T
    _close(dx.reshape(T, B, I), xt.grad.numpy(), 1e-2, "dx")
    L.check(lib.ocr_birnn_set_path(0), "path")


@pytest.mark.parametrize("B,cell,sizes", [(4, "lstm", (32, 32)), (3, "lstm", (32, 32)), (4, "gru", (32, 16)), (3, "gru", (32, 16))])
def test_train_step_vs_oracle(B, cell, sizes):
    """One full step of train.py's graph: loss, every gradient, moving statistics and the Adam update
    (LSTM model of model_bu.py and GRU model of model.py)."""
    from cnn_lstm_ctc_ocr_b200 import train
    from oracle import train_oracle as to
    params, img, widths, labels = _small_problem(B=B, cell=cell, sizes=sizes)
    ref = to.train_step_reference(params, img, widths, labels, step=0, cell_type=cell, sizes=sizes)
    tr = train.Trainer(params, cell_type=cell, rnn_sizes=sizes)
    losses = tr.forward_backward(torch.tensor(img, device=DEV), widths, labels)
    _close(losses.cpu().numpy(), ref["losses"], 5e-3, "losses")
    _close(tr.last_logits.cpu().numpy(), ref["logits"], 1e-2, "logits")
    # Recurrent layers and logits: smooth functions of their inputs -> max-norm parity.  Convolutional stack: eight
    # ReLU / max-pool stages make the gradient a DISCONTINUOUS function of the activations; TF32-level noise (3e-4) flips a
    # fraction of the masks and moves the float64 oracle's own conv gradients by ~10% in L2
    # (tests/test_train_oracle.py::test_conv_gradient_conditioning measures it), so these are asserted in L2 / direction;
    # their element-wise parity is asserted op by op above, on identical inputs.
    bad = {}
    for name, g in ref["grads"].items():
        got = tr.grads[name].cpu().numpy().astype(np.float64)
        scale = np.abs(g).max()
        if scale < 1e-12:          # conv biases in front of a batch-norm: the true gradient is exactly zero
            assert np.abs(got).max() < 1e-5, name
            continue
        if name.startswith("rnn/"):
            e = np.abs(got - g).max() / scale
            if e > 3e-2:
                bad[name] = ("max", e)
        else:
            l2 = np.linalg.norm(got - g) / np.linalg.norm(g)
            cos = float((got * g).sum() / (np.linalg.norm(got) * np.linalg.norm(g)))
            if l2 > 0.3 or cos < 0.95:
                bad[name] = ("l2", l2, "cos", cos)
    assert not bad, "gradients off: %s" % bad
    for name, v in tr.stats.items():
        _close(v.cpu().numpy(), ref["new_params"][name], 1e-3, name)
    loss = float(losses.mean())
    tr.apply_gradients()
    assert tr.global_step == 1
    # Adam's first step moves every weight by lr * g/(|g| + eps'), i.e. by the SIGN of its gradient: compare the smooth
    # (recurrent / logits) variables where the gradient is well away from zero; ocr_adam_step itself is checked above
    for name, g in ref["grads"].items():
        if not name.startswith("rnn/"):
            continue
        big = np.abs(g) > 5e-2 * np.abs(g).max()
        if big.any() and np.abs(g).max() > 1e-9:
            got = tr.params[name].cpu().numpy()
            assert np.abs(got - ref["new_params"][name])[big].max() < 2e-6, name
    # and training makes progress: a few more steps on the same batch lower the loss
    for _ in range(25):
        last = float(tr.train_step(torch.tensor(img, device=DEV), widths, labels))
    assert np.isfinite(last) and last < loss
    m = tr.to_model()
    assert len(m.recognize(torch.tensor(img, device=DEV), torch.tensor(widths))) == len(labels)


def test_pool_arg_step_matches_separate_pool_step():
    """The whole backward pass with the pooled layers on the arguments-of-the-maxima path against the same pass with the
    full-size activations and ocr_maxpool_bwd: same losses bit for bit (the forward bits are the same), gradients equal up to the
    summation order of the batch-norm sums (float64 accumulators of float partials)."""
    from cnn_lstm_ctc_ocr_b200 import train
    params, img, widths, labels = _small_problem(B=4, cell="lstm", sizes=(32, 32))
    out = []
    for flag in (True, False):
        tr = train.Trainer(params, cell_type="lstm", rnn_sizes=(32, 32))
        tr.pool_arg = flag
        losses = tr.forward_backward(torch.tensor(img, device=DEV), widths, labels)
        torch.cuda.synchronize()
        out.append((losses.cpu().numpy(), {k: v.cpu().numpy().copy() for k, v in tr.grads.items()}))
    assert np.array_equal(out[0][0], out[1][0])
    for k, g in out[0][1].items():
        ref = out[1][1][k]
        assert np.abs(g - ref).max() <= 2e-5 * (np.abs(ref).max() + 1e-12), k


def test_captured_step_matches_eager_step():
    """The CUDA-graph form of the step (Trainer.capture / train_step_captured) replays the same kernels: same losses,
    same updated variables as the eager step, step after step (the learning rate reaches Adam through device memory)."""
    from cnn_lstm_ctc_ocr_b200 import train
    params, img, widths, labels = _small_problem(B=4)
    a = train.Trainer(params, rnn_sizes=(32, 32))
    b = train.Trainer(params, rnn_sizes=(32, 32))
    b.capture(4, img.shape[2], max_label_len=8)
    dimg = torch.tensor(img, device=DEV)
    for step in range(3):
        la = a.forward_backward(dimg, widths, labels)
        a.apply_gradients()
        lb = b.train_step_captured(dimg, widths, labels)
        _close(lb.cpu().numpy(), la.cpu().numpy(), 1e-5, "losses step %d" % step)
    assert a.global_step == b.global_step == 3
    _close(b.theta.cpu().numpy(), a.theta.cpu().numpy(), 1e-5, "variables")
    for k in a.stats:
        _close(b.stats[k].cpu().numpy(), a.stats[k].cpu().numpy(), 1e-5, k)


def test_data_parallel_step():
    """2 GPUs: batch-sharded step with the NCCL gradient all-reduce == single-replica step (tests/ddp_train_check.py)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "tests", "ddp_train_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DDP_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
