"""GPU parity: proposal generation, fused instance cost volume (+gate), soft-argmin -- vs the oracle and goldens."""
import hashlib

import numpy as np
import pytest
import torch

from conftest import golden, rel_err

pytestmark = pytest.mark.gpu

from oracle import c_oracle as co  # noqa: E402
from oracle import torch_port as tp  # noqa: E402


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("D", [16, 48])
def test_proposal_shift_golden_bit_exact(lib, D):
    from side_b200.networks import get_proposal_shift
    g = golden("proposal_shift_D%d" % D)
    pl, pr, db = get_proposal_shift(dev(g["left"]), dev(g["right"]), D, dev(g["fb"]), None)
    assert np.array_equal(db.cpu().numpy(), g["depth_bin"])
    assert np.array_equal(pl.cpu().numpy(), g["pro_left"])
    assert np.array_equal(pr.cpu().numpy(), g["pro_right"])


def test_inst_costvol_golden(lib):
    """The reference loop stereo_network_old.py:366-376 + gate :197-203 + tail :228-236 (golden from the reference)."""
    from side_b200 import ops
    g = golden("inst_costvol")
    fL, fR = dev(g["featL"].astype(np.float32)), dev(g["featR"].astype(np.float32))
    left, right, fb = dev(g["left"]), dev(g["right"]), dev(g["fb"])
    cost, db = ops.inst_costvol(fL, fR, left, right, fb, 16, 16, 319.0)
    c = cost.cpu().numpy()
    assert hashlib.sha256(c.tobytes()).hexdigest() == str(g["cost_sha256"]), "raw [L,R,L-R] volume must be bit-exact"
    assert np.array_equal(db.cpu().numpy(), g["depth_bin"])
    gated, _ = ops.inst_costvol(fL, fR, left, right, fb, 16, 16, 319.0, gate=True)
    assert rel_err(gated.cpu().numpy().reshape(-1)[::97], g["gated_sample"]) < 1e-4
    gated2 = ops.xcross_gate(cost, 32)                       # stand-alone gate == fused gate
    assert rel_err(gated2.cpu().numpy(), gated.cpu().numpy()) < 1e-5
    depth = ops.softargmin(dev(g["logits"][:, 0]), db)
    assert rel_err(depth.cpu().numpy(), g["disp"]) < 1e-4   # north_star: 1e-4 relative for soft-argmin depth


def _random_case(rng, B, C, H, W, N):
    fL = rng.standard_normal((B, C, H, W)).astype(np.float32)
    fR = rng.standard_normal((B, C, H, W)).astype(np.float32)
    b = np.sort(rng.integers(0, B, N)).astype(np.float32)
    x1 = rng.uniform(-6, W - 4, N); w = rng.uniform(0.2, W / 3, N)
    y1 = rng.uniform(-4, H - 3, N); h = rng.uniform(0.2, H / 2, N)
    sh = rng.uniform(0, 12, N)
    left = np.stack([b, x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    right = np.stack([b, x1 - sh, y1 + rng.uniform(-1, 1, N), x1 + w - sh, y1 + h], 1).astype(np.float32)
    fb = rng.uniform(300, 450, B).astype(np.float32)
    return fL, fR, left, right, fb


@pytest.mark.parametrize("cfg", [(2, 8, 24, 80, 9, 16, 16), (1, 3, 10, 33, 4, 2, 7), (3, 32, 17, 61, 11, 5, 16), (1, 64, 24, 80, 3, 48, 16)])
def test_inst_costvol_random_bit_exact(lib, cfg):
    """Boxes partly outside the image, sub-pixel boxes, P != 16, C not a power of two, D = 2 .. 48."""
    from side_b200 import ops
    B, C, H, W, N, D, P = cfg
    rng = np.random.default_rng(B * 100 + C)
    fL, fR, left, right, fb = _random_case(rng, B, C, H, W, N)
    pl, pr, db = co.proposal_shift(left, right, fb, D, x_clamp=W - 1.0)
    ref = co.inst_costvol(fL, fR, pl, pr, P)
    cost, dbin = ops.inst_costvol(dev(fL), dev(fR), dev(left), dev(right), dev(fb), D, P, W - 1.0)
    c = cost.cpu().numpy()
    assert np.array_equal(dbin.cpu().numpy(), db)
    assert np.array_equal(c, ref), "max abs diff %g" % np.abs(c - ref).max()
    assert np.array_equal(c[:, 2 * C:], c[:, :C] - c[:, C:2 * C])       # L-R is computed from the kernel's own L, R
    gref, xref = co.xcross_gate(ref, C)
    gated, _ = ops.inst_costvol(dev(fL), dev(fR), dev(left), dev(right), dev(fb), D, P, W - 1.0, gate=True)
    assert rel_err(gated.cpu().numpy(), gref) < 1e-4          # north_star: 1e-4 relative for the correlation/gate


def test_inst_costvol_valid_mask_and_empty(lib):
    from side_b200 import ops
    rng = np.random.default_rng(1)
    fL, fR, left, right, fb = _random_case(rng, 2, 8, 12, 40, 6)
    valid = np.array([1, 0, 1, 1, 0, 1], np.uint8)
    cost, db = ops.inst_costvol(dev(fL), dev(fR), dev(left), dev(right), dev(fb), 4, 16, 39.0, gate=True, valid=dev(valid))
    full, dbf = ops.inst_costvol(dev(fL), dev(fR), dev(left), dev(right), dev(fb), 4, 16, 39.0, gate=True)
    c, f = cost.cpu().numpy(), full.cpu().numpy()
    assert np.all(c[valid == 0] == 0) and np.all(db.cpu().numpy()[valid == 0] == 0)
    assert np.array_equal(c[valid == 1], f[valid == 1])
    e, edb = ops.inst_costvol(dev(fL), dev(fR), dev(left[:0]), dev(right[:0]), dev(fb), 4, 16, 39.0)
    assert tuple(e.shape) == (0, 24, 4, 16, 16) and tuple(edb.shape) == (0, 4)


@pytest.mark.parametrize("gate", [False, True])
def test_inst_costvol_backward_vs_autograd_of_port(lib, gate):
    """Gradients w.r.t. the feature maps vs torch.autograd through the reference-style RoIAlign loop (CPU)."""
    from side_b200 import ops
    rng = np.random.default_rng(7)
    fL, fR, left, right, fb = _random_case(rng, 2, 6, 14, 44, 5)
    D, P = 5, 16
    tl = [torch.from_numpy(a).requires_grad_(True) for a in (fL, fR)]
    cref, _ = tp.inst_costvol(tl[0], tl[1], torch.from_numpy(left), torch.from_numpy(right), torch.from_numpy(fb), D, P, 43.0, gate=gate)
    gc = torch.from_numpy(rng.standard_normal(tuple(cref.shape)).astype(np.float32))
    gref = torch.autograd.grad(cref, tl, gc)
    tg = [dev(a).requires_grad_(True) for a in (fL, fR)]
    cost, _ = ops.inst_costvol(tg[0], tg[1], dev(left), dev(right), dev(fb), D, P, 43.0, gate=gate)
    g = torch.autograd.grad(cost, tg, gc.cuda())
    for mine, r, n in zip(g, gref, ("gfeatL", "gfeatR")):
        assert rel_err(mine.cpu().numpy(), r.numpy()) < 1e-4, n


def test_gate_and_softargmin_backward(lib):
    from side_b200 import ops
    torch.manual_seed(2)
    cost = torch.randn(3, 12, 5, 4, 4)
    a = cost.clone().requires_grad_(True)
    ref = tp.xcross_gate(a, 4)
    go = torch.randn_like(ref)
    gr, = torch.autograd.grad(ref, a, go)
    b = cost.cuda().requires_grad_(True)
    out = ops.xcross_gate(b, 4)
    gm, = torch.autograd.grad(out, b, go.cuda())
    assert rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
    assert rel_err(gm.cpu().numpy(), gr.numpy()) < 1e-4
    for (N, D) in [(7, 16), (64, 48), (3, 200), (1, 1)]:
        lg = torch.randn(N, D, 4, 4); db = torch.rand(N, D) * 80 + 1
        l1 = lg.clone().requires_grad_(True); d1 = db.clone().requires_grad_(True)
        r = tp.softargmin(l1, d1)
        gd = torch.randn(N)
        grl, grd = torch.autograd.grad(r, (l1, d1), gd)
        l2 = lg.cuda().requires_grad_(True); d2 = db.cuda().requires_grad_(True)
        o = ops.softargmin(l2, d2)
        gl, gdb = torch.autograd.grad(o, (l2, d2), gd.cuda())
        dref, _ = co.softargmin(lg.numpy(), db.numpy())
        assert rel_err(o.detach().cpu().numpy(), dref) < 1e-5, (N, D)
        assert rel_err(gl.cpu().numpy(), grl.numpy()) < 1e-4 and rel_err(gdb.cpu().numpy(), grd.numpy()) < 1e-4, (N, D)


def test_config2_full_size_properties(lib):
    """BASELINE config #2: 64 RoIs, 48 candidates, 64 channels (604 MB volume).  Size-independent checks:
    a seeded subset of RoIs against the oracle (bit-exact), L-R identity on the whole volume, gate scaling."""
    from side_b200 import ops
    from side_b200.utils.synthetic import make_boxes
    torch.manual_seed(0)
    fL, fR = torch.randn(1, 64, 96, 320), torch.randn(1, 64, 96, 320)
    left, right, _ = make_boxes(1, 64, seed=0)
    fb = torch.tensor([384.38])
    cost, db = ops.inst_costvol(fL.cuda(), fR.cuda(), left.cuda(), right.cuda(), fb.cuda(), 48, 16, 319.0)
    assert tuple(cost.shape) == (64, 192, 48, 16, 16)
    assert torch.equal(cost[:, 128:], cost[:, :64] - cost[:, 64:128])
    sub = [0, 17, 63]
    pl, pr, dbo = co.proposal_shift(left.numpy()[sub], right.numpy()[sub], fb.numpy(), 48)
    ref = co.inst_costvol(fL.numpy(), fR.numpy(), pl, pr, 16)
    assert np.array_equal(cost[sub].cpu().numpy(), ref)
    assert np.array_equal(db[sub].cpu().numpy(), dbo)
    gated, _ = ops.inst_costvol(fL.cuda(), fR.cuda(), left.cuda(), right.cuda(), fb.cuda(), 48, 16, 319.0, gate=True)
    ratio = (gated[:, :64] * cost[:, :64]).sum((1, 3, 4)) / (cost[:, :64] ** 2).sum((1, 3, 4))
    l, r = cost[:, :64].double(), cost[:, 64:128].double()
    xc = (l * r).sum((1, 3, 4)) / torch.clamp(torch.sqrt((l * l).sum((1, 3, 4))) * torch.sqrt((r * r).sum((1, 3, 4))), min=0.01)
    assert (ratio.double() - xc).abs().max().item() < 1e-5


@pytest.mark.parametrize("cfg", [(2, 8, 24, 80, 9, 16, 16), (1, 64, 24, 80, 3, 48, 16), (3, 32, 17, 61, 11, 5, 16), (1, 4, 10, 33, 4, 2, 6)])
def test_nchw_and_nhwc_gather_paths_identical(lib, cfg):
    """The channels-last gather (default) and the NCHW gather produce the same bits, with and without the gate."""
    from side_b200 import ops
    B, C, H, W, N, D, P = cfg
    rng = np.random.default_rng(C + N)
    fL, fR, left, right, fb = _random_case(rng, B, C, H, W, N)
    args = (dev(fL), dev(fR), dev(left), dev(right), dev(fb), D, P, W - 1.0)
    res = {}
    try:
        for mode in (True, False):
            ops.USE_NHWC_GATHER = mode
            res[mode] = (ops.inst_costvol(*args)[0], ops.inst_costvol(*args, gate=True)[0])
    finally:
        ops.USE_NHWC_GATHER = True
    assert torch.equal(res[True][0], res[False][0])
    assert (res[True][1] - res[False][1]).abs().max().item() <= 1e-6 * res[False][1].abs().max().item()


def test_fma_mode_within_1e6(lib):
    """SIDE_VOL_FMA (opt-in): FMA-contracted taps; <= 1e-6 of the volume's range from the bit-exact default."""
    from side_b200 import ops
    rng = np.random.default_rng(11)
    fL, fR, left, right, fb = _random_case(rng, 2, 32, 24, 80, 9)
    args = (dev(fL), dev(fR), dev(left), dev(right), dev(fb), 16, 16, 79.0)
    for gate in (False, True):
        a = ops.inst_costvol(*args, gate=gate)[0]
        b = ops.inst_costvol(*args, gate=gate, fma=True)[0]
        assert (a - b).abs().max().item() <= 1e-6 * a.abs().max().item()


@pytest.mark.parametrize("cfg", [(2, 8, 24, 80, 9, 16), (1, 64, 24, 80, 3, 48), (3, 32, 17, 61, 11, 5), (1, 16, 96, 320, 7, 48)])
def test_separable_path_vs_exact_oracle(lib, cfg):
    """SIDE_VOL_SEPARABLE: same sample positions / validity rules, products re-associated (y first, shared by all D).
    Interpolated values <= 1e-5 of the range from the bit-exact oracle (SURVEY 8(a) A5 bound); L-R from the kernel's own
    L, R exactly; depth bins bit-exact; gate <= 1e-4; ungated + xcross single pass == gated."""
    from side_b200 import ops
    B, C, H, W, N, D = cfg
    rng = np.random.default_rng(B * 7 + C + D)
    fL, fR, left, right, fb = _random_case(rng, B, C, H, W, N)
    if W == 320:   # boxes clamped at the right border, at 0, a sub-pixel box and one wider than the 96-column window
        left[:4, 1:] = [[300., 5., 318., 40.], [2., 10., 30., 20.], [100.2, 50.1, 100.6, 50.4], [40., 3., 290., 90.]]
        right[:4, 1:] = [[290., 5., 309., 41.], [-6., 10., 22., 20.], [97.2, 50.1, 97.7, 50.4], [20., 3., 270., 90.]]
    pl, pr, db = co.proposal_shift(left, right, fb, D, x_clamp=W - 1.0)
    ref = co.inst_costvol(fL, fR, pl, pr, 16)
    args = (dev(fL), dev(fR), dev(left), dev(right), dev(fb), D, 16, W - 1.0)
    cost, dbin = ops.inst_costvol(*args, separable=True)
    c = cost.cpu().numpy()
    assert np.array_equal(dbin.cpu().numpy(), db)
    assert rel_err(c, ref) < 1e-5
    assert np.array_equal(c[:, 2 * C:], c[:, :C] - c[:, C:2 * C])
    gref, xref = co.xcross_gate(ref, C)
    gated, _ = ops.inst_costvol(*args, gate=True, separable=True)
    assert rel_err(gated.cpu().numpy(), gref) < 1e-4
    raw, db2, xc = ops.inst_costvol_ungated(*args)
    assert torch.equal(raw, cost) and torch.equal(db2, dbin)
    assert rel_err(xc.cpu().numpy(), xref) < 1e-4
    assert rel_err((raw * xc[:, None, :, None, None]).cpu().numpy(), gated.cpu().numpy()) < 1e-6
    # deterministic: a second run gives the same bits
    assert torch.equal(ops.inst_costvol(*args, gate=True, separable=True)[0], gated)


def test_separable_path_valid_mask_and_absurd_boxes(lib):
    from side_b200 import ops
    rng = np.random.default_rng(5)
    fL, fR, left, right, fb = _random_case(rng, 2, 8, 12, 40, 6)
    left[2, 1:] = [-3000., 2., 39., 9.]          # > 1500 columns wide: slow path inside the kernel
    right[2, 1:] = [-3010., 2., 30., 9.]
    left[3, 1:] = [50., 2., 60., 9.]             # completely right of the 40-column image
    right[3, 1:] = [45., 2., 55., 9.]
    valid = np.array([1, 0, 1, 1, 0, 1], np.uint8)
    args = (dev(fL), dev(fR), dev(left), dev(right), dev(fb), 4, 16, 39.0)
    pl, pr, db = co.proposal_shift(left, right, fb, 4, x_clamp=39.0)
    ref = co.inst_costvol(fL, fR, pl, pr, 16)
    full, _ = ops.inst_costvol(*args, separable=True)
    assert rel_err(full.cpu().numpy(), ref) < 1e-5
    gref, _ = co.xcross_gate(ref, 8)
    cost, dbm = ops.inst_costvol(*args, gate=True, valid=dev(valid), separable=True)
    c = cost.cpu().numpy()
    assert np.all(c[valid == 0] == 0) and np.all(dbm.cpu().numpy()[valid == 0] == 0)
    assert rel_err(c[valid == 1], gref[valid == 1]) < 1e-4
    raw, _, xc = ops.inst_costvol_ungated(*args, valid=dev(valid))
    assert np.all(raw.cpu().numpy()[valid == 0] == 0) and np.all(xc.cpu().numpy()[valid == 0] == 0)


# ------------------------------------------------------------------------------------------------
# separable gather backward (inst_costvol_bwd_gather_kernel): P == 16, C % 8 == 0
# ------------------------------------------------------------------------------------------------
def _grads(ops, args, gc, gate, valid=None):
    tg = [a.clone().requires_grad_(True) for a in args[:2]]
    cost, _ = ops.inst_costvol(tg[0], tg[1], *args[2:], gate=gate, valid=valid)
    return torch.autograd.grad(cost, tg, gc)


@pytest.mark.parametrize("gate", [False, True])
@pytest.mark.parametrize("cfg", [(2, 8, 14, 44, 5, 5), (1, 16, 24, 80, 6, 48), (1, 8, 96, 320, 7, 16)])
def test_gather_backward_vs_scalar_kernel_and_port(lib, cfg, gate):
    """The register-accumulating gather backward against the scalar-atomic kernel (itself checked against autograd of
    the reference-style RoIAlign loop), on boxes at both borders, a sub-pixel box, a box wider than one 192-column
    pass, an invalid RoI -- and directly against the CPU autograd of the port for the small case."""
    from side_b200 import _lib, ops
    B, C, H, W, N, D = cfg
    rng = np.random.default_rng(C + D + W)
    fL, fR, left, right, fb = _random_case(rng, B, C, H, W, N)
    if W == 320:
        left[:4, 1:] = [[300., 5., 318., 40.], [2., 10., 30., 20.], [100.2, 50.1, 100.6, 50.4], [40., 3., 290., 90.]]
        right[:4, 1:] = [[290., 5., 309., 41.], [-6., 10., 22., 20.], [97.2, 50.1, 97.7, 50.4], [20., 3., 270., 90.]]
    valid = np.ones(N, np.uint8)
    valid[N - 1] = 0
    args = (dev(fL), dev(fR), dev(left), dev(right), dev(fb), D, 16, W - 1.0)
    gc = torch.from_numpy(rng.standard_normal((N, 3 * C, D, 16, 16)).astype(np.float32)).cuda()
    g = _grads(ops, args, gc, gate, dev(valid))
    ops.VOL_BWD_FLAGS = _lib.VOL_BWD_SCALAR
    try:
        gs = _grads(ops, args, gc, gate, dev(valid))
    finally:
        ops.VOL_BWD_FLAGS = 0
    for mine, r, n in zip(g, gs, ("gfeatL", "gfeatR")):
        assert rel_err(mine.cpu().numpy(), r.cpu().numpy()) < 1e-4, (n, cfg, gate)
    if W == 44:
        tl = [torch.from_numpy(a).requires_grad_(True) for a in (fL, fR)]
        keep = valid.astype(bool)
        cref, _ = tp.inst_costvol(tl[0], tl[1], torch.from_numpy(left[keep]), torch.from_numpy(right[keep]), torch.from_numpy(fb),
                                  D, 16, W - 1.0, gate=gate)
        gref = torch.autograd.grad(cref, tl, gc.cpu()[torch.from_numpy(keep)])
        for mine, r, n in zip(g, gref, ("gfeatL", "gfeatR")):
            assert rel_err(mine.cpu().numpy(), r.numpy()) < 1e-4, ("port", n, gate)


def test_gather_backward_config2_linearity(lib):
    """config #2 size (64 RoIs x 48 candidates x 64 channels): <g, J v> == <J^T g, v> -- the backward is the exact
    adjoint of the (linear, ungated) forward -- and the scalar kernel agrees."""
    from side_b200 import _lib, ops
    from side_b200.utils.synthetic import make_boxes
    torch.manual_seed(1)
    fL, fR = torch.randn(1, 64, 96, 320, device="cuda"), torch.randn(1, 64, 96, 320, device="cuda")
    left, right, _ = make_boxes(1, 64, seed=0)
    fb = torch.tensor([384.38], device="cuda")
    args = (fL, fR, left.cuda(), right.cuda(), fb, 48, 16, 319.0)
    cost, _ = ops.inst_costvol(*args)
    gc = torch.randn_like(cost)
    lhs = (cost.double() * gc.double()).sum().item()
    scale = (cost.double().norm() * gc.double().norm()).item()
    del cost
    g = _grads(ops, args, gc, False)
    rhs = ((g[0].double() * fL.double()).sum() + (g[1].double() * fR.double()).sum()).item()
    assert abs(lhs - rhs) <= 1e-5 * scale, (lhs, rhs, scale)
    ops.VOL_BWD_FLAGS = _lib.VOL_BWD_SCALAR
    try:
        gs = _grads(ops, args, gc, False)
    finally:
        ops.VOL_BWD_FLAGS = 0
    for a, b2 in zip(g, gs):
        assert rel_err(a.cpu().numpy(), b2.cpu().numpy()) < 1e-4


# ------------------------------------------------------------------------------------------------
# volume emitted directly as gated channels-last fp16 pairs (inst_costvol_cl.cu)
# ------------------------------------------------------------------------------------------------
def _cl_reference(fL, fR, left, right, fb, D, x_clamp, gate):
    """Oracle: bit-exact RoIAlign volume (+ gate) in the reference layout, permuted to [N, D, 16, 16, 3C]."""
    pl, pr, db = co.proposal_shift(left, right, fb, D, x_clamp=x_clamp)
    raw = co.inst_costvol(fL, fR, pl, pr, 16)
    gated, xc = co.xcross_gate(raw, fL.shape[1])
    vol = gated if gate else raw
    return np.ascontiguousarray(vol.transpose(0, 2, 3, 4, 1)), db, xc


@pytest.mark.parametrize("case", ["typical", "wide", "garbage", "D48", "ungated"])
def test_volume_channels_last_pairs_vs_oracle(lib, case):
    """hi + lo / 2^11 of side_inst_costvol_fwd_cl against the bit-exact oracle volume (gate applied): <= 1e-5 of the range for
    the separable evaluation order plus the 2^-21 resolution of the pair; depth bins bit-exact; gate scalars <= 1e-5;
    rows with valid == 0 are all-zero.  'wide' / 'garbage' exercise the column-tiled two-pass mode (boxes wider than the
    40-column window), clamping at both image borders and degenerate boxes."""
    from side_b200 import ops
    rng = np.random.default_rng({"typical": 1, "wide": 2, "garbage": 3, "D48": 4, "ungated": 5}[case])
    B, C, H, W = 2, 32, 96, 320
    D = 48 if case == "D48" else 16
    N = 24
    fL = rng.standard_normal((B, C, H, W)).astype(np.float32)
    fR = rng.standard_normal((B, C, H, W)).astype(np.float32)
    b = np.sort(rng.integers(0, B, N)).astype(np.float32)
    if case in ("typical", "D48", "ungated"):
        x1 = rng.uniform(20, 270, N); w = rng.uniform(8, 36, N); y1 = rng.uniform(5, 70, N); h = rng.uniform(6, 25, N)
    elif case == "wide":
        x1 = rng.uniform(0, 120, N); w = rng.uniform(45, 300, N); y1 = rng.uniform(0, 40, N); h = rng.uniform(10, 90, N)
    else:
        x1 = rng.uniform(-40, 330, N); w = rng.uniform(-5, 400, N); y1 = rng.uniform(-20, 100, N); h = rng.uniform(-3, 120, N)
        w[:3] = 0.0; h[3:5] = 0.0
    sh = rng.uniform(0, 14, N)
    left = np.stack([b, x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
    right = np.stack([b, x1 - sh, y1 + rng.uniform(-1, 1, N), x1 + w - sh, y1 + h], 1).astype(np.float32)
    fb = rng.uniform(300, 450, B).astype(np.float32)
    valid = np.ones(N, np.uint8)
    valid[[2, 11]] = 0
    gate = case != "ungated"
    ref, db, xc = _cl_reference(fL, fR, left, right, fb, D, 319.0, gate)
    hi, lo, dbin, xcr = ops.inst_costvol_cl(dev(fL), dev(fR), dev(left), dev(right), dev(fb), D, 16, 319.0, valid=dev(valid), gate=gate)
    assert tuple(hi.shape) == (N, D, 16, 16, 96) and hi.dtype == torch.float16
    rec = (hi.float() + lo.float() / 2048.0).cpu().numpy()
    live = valid.astype(bool)
    assert np.array_equal(dbin.cpu().numpy()[live], db[live])
    assert not rec[~live].any() and not dbin.cpu().numpy()[~live].any()
    scale = np.abs(ref[live]).max()
    assert np.abs(rec[live] - ref[live]).max() <= 1.2e-5 * scale, case
    assert np.abs(xcr.cpu().numpy()[live] - xc[live]).max() <= 1e-5, case
    # L - R plane: exactly (L - R) * g of the kernel's own L and R is not observable after the split; check consistency to pair resolution
    l, r, dd = rec[live][..., :32], rec[live][..., 32:64], rec[live][..., 64:]
    assert np.abs(dd - (l - r)).max() <= 4e-6 * scale


def test_volume_channels_last_pairs_match_two_pass_path(lib):
    """Same inputs through the round-1 route (separable NCDHW volume + xcross, then side_ncdhw_to_cl_split_f16 with the gate as
    scale): the fused kernel reproduces its pairs (identical arithmetic order up to the gate's reduction order), and the
    aggregation network fed with either gives the same depth."""
    from side_b200 import ops
    from side_b200.networks.stereo_network import cost_volume
    from side_b200.utils.synthetic import make_boxes
    torch.manual_seed(3)
    fL, fR = torch.randn(2, 32, 96, 320, device="cuda"), torch.randn(2, 32, 96, 320, device="cuda")
    left, right, _ = make_boxes(2, 20, seed=4)
    left, right = left.cuda(), right.cuda()
    fb = torch.tensor([384.38, 400.0], device="cuda")
    cost, db0, xc0 = ops.inst_costvol_ungated(fL, fR, left, right, fb, 16, 16, 319.0)
    hi0, lo0 = ops.ncdhw_to_cl_split(cost, scale=xc0, fmt="f16")
    hi1, lo1, db1, xc1 = ops.inst_costvol_cl(fL, fR, left, right, fb, 16, 16, 319.0)
    assert torch.equal(db0, db1)
    assert float((xc0 - xc1).abs().max()) <= 2e-6
    a = hi0.float() + lo0.float() / 2048.0
    c = hi1.float() + lo1.float() / 2048.0
    assert float((a - c).abs().max()) <= 3e-6 * float(a.abs().max())
    m = cost_volume(64).cuda().eval()
    with torch.no_grad():
        d0 = ops.softargmin(m.aggregate_tc_pairs(hi0, lo0, "f16"), db0)
        d1 = ops.softargmin(m.aggregate_tc_pairs(hi1, lo1, "f16"), db1)
    assert float(((d0 - d1).abs() / d0.abs()).max()) < 1e-5


@pytest.mark.gpu
def test_volume_without_difference_plane_and_folded_first_convolution(lib):
    """SIDE_VOL_NO_DIFF: the builder emits only the (L, R) planes -- bit-identical to the first 2C channels of the full volume --
    and cost_volume.dres0.0 with the L - R plane folded into its weights (W_L + W_D, W_R - W_D; the volume is cat(L, R, L - R),
    stereo_network_old.py:374-376, and this convolution is its only reader) gives the same logits and depth: <= 1e-5 of the logit
    range, depth <= 1e-5 relative.  Also against the module's own fp64 convolution stack on the full volume."""
    from side_b200 import ops
    from side_b200.networks.stereo_network import cost_volume
    from side_b200.utils.synthetic import make_boxes
    torch.manual_seed(5)
    fL, fR = torch.randn(2, 32, 96, 320, device="cuda"), torch.randn(2, 32, 96, 320, device="cuda")
    left, right, _ = make_boxes(2, 12, seed=6)
    left, right = left.cuda(), right.cuda()
    fb = torch.tensor([384.38, 400.0], device="cuda")
    valid = torch.ones(24, dtype=torch.uint8, device="cuda")
    valid[5] = 0
    hi3, lo3, db3, xc3 = ops.inst_costvol_cl(fL, fR, left, right, fb, 16, 16, 319.0, valid=valid)
    hi2, lo2, db2, xc2 = ops.inst_costvol_cl(fL, fR, left, right, fb, 16, 16, 319.0, valid=valid, diff=False)
    assert tuple(hi2.shape) == (24, 16, 16, 16, 64)
    assert torch.equal(hi2, hi3[..., :64]) and torch.equal(lo2, lo3[..., :64])
    assert torch.equal(db2, db3) and torch.equal(xc2, xc3)
    m = cost_volume(64).cuda().eval()
    with torch.no_grad():
        for b in m.modules():                        # non-trivial folded BatchNorm
            if isinstance(b, (torch.nn.BatchNorm3d, torch.nn.BatchNorm2d)):
                b.running_mean.normal_(0, 0.1); b.running_var.uniform_(0.5, 1.5); b.weight.uniform_(0.5, 1.5); b.bias.normal_(0, 0.1)
        l3 = m.aggregate_tc_pairs(hi3, lo3, "f16")
        l2 = m.aggregate_tc_pairs(hi2, lo2, "f16", folded=True)
        assert float((l3 - l2).abs().max()) <= 1e-5 * float(l3.abs().max())
        keep = valid.bool()
        d3, d2 = ops.softargmin(l3, db3)[keep], ops.softargmin(l2, db2)[keep]
        assert float(((d3 - d2).abs() / d3.abs()).max()) < 1e-5
        # fp64 module on the full [L, R, L - R] volume (NCDHW)
        import copy
        vol = (hi3.double() + lo3.double() / 2048.0).permute(0, 4, 1, 2, 3).contiguous().cpu()
        ref = copy.deepcopy(m).cpu().double().aggregate(vol)
        assert float((l2.double().cpu() - ref).abs().max()) <= 1e-4 * float(ref.abs().max())
