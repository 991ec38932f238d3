"""CPU: the reference-facing Python API (NODE_WAN_solver / loss / func_eval / Comb_loader /
autograd.Function) driven end to end with the CPU EMULATION build of the kernels injected, against
the golden vectors of the unmodified reference.  Checks host logic only (packing, strides,
normalisers, side-effect terms, state-dict names, sampler RNG order); the real sm_100a library is
checked by the `-m gpu` tests."""
import os

import numpy as np
import pytest
import torch

import xnode_wan_b200 as xw
from tests import _golden as G
from tests.host_emu import build_emu


@pytest.fixture()
def emu(monkeypatch):
    lib = xw._lib.XwLib(build_emu.build())
    monkeypatch.setattr(xw._lib, "_LIB", lib)
    return lib


def make_solver(case, device="cpu"):
    p = dict(case["params"])
    p["domain"] = case["meta"].get("domain_class", "Hypercube")
    prob = xw.problems.by_name(case["meta"]["funcs"], p["dim"])
    if "coef_a" in case["z"].files:          # the constant a != I this golden was made with
        A = case["z"]["coef_a"]
        prob.func_a = lambda X_, i, j: torch.full(X_.shape[:-1], float(A[i, j]))
    if case["meta"].get("general_coef"):     # a_ij(X) varying over the sample, c(X, u) non-affine
        from tests import _general_coef as GC
        prob.func_a, prob.func_c = GC.func_a, GC.func_c
    s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, device,
                           "./", func_u_sol=prob.func_u_sol, p=2, log_json=False)
    with torch.no_grad():
        for q, w in zip(s.u_net.parameters(), case["thu_list"]):
            q.copy_(torch.from_numpy(np.asarray(w)))
        for q, w in zip(s.v_net.parameters(), case["thv_list"]):
            q.copy_(torch.from_numpy(np.asarray(w)))
    return s, prob


def eval_phase(s, case, phase, device="cpu"):
    z = case["z"]
    X, XV, BX = (torch.from_numpy(z[k]).to(device) for k in ("X", "XV", "BX"))
    dom = s.new_domain()
    s.optimizer_u.zero_grad()
    s.optimizer_v.zero_grad()
    pv, pu = s.v_net(XV), s.u_net(X)
    h, f, g, a, b, c = xw.func_eval(X, BX, s.setup, pu, s.func_a, s.func_b, s.func_c, s.func_h, s.func_f, s.func_g)
    L = xw.loss(s.config["alpha"], a, b, c, h, f, g, s.setup, dom, device)
    val = L.u(pu, pv, s.u_net, X, XV, BX) if phase == "u" else L.v(pu, pv, X, XV)
    val.backward()
    net = s.u_net if phase == "u" else s.v_net
    return val, [(q.grad if q.grad is not None else torch.zeros_like(q)).detach().cpu().numpy() for q in net.parameters()]


@pytest.mark.parametrize("name", ["cube_d5_alpha1_randbias", "cube_d3_small_nets", "cube_d4_ex43", "cube_d3_rk4",
                                  "cube_d4_aconst", "cone_d5_g2", "hourglass_d5_g2_reentry", "hourglass_d5_g18",
                                  "extra/general_coef_cube_d3"])
def test_reference_api_matches_golden(emu, name):
    case = G.load(name)
    s, _ = make_solver(case)
    z = case["z"]
    for phase, gold_l, gold_g in (("u", float(z["loss_u"]), case["gu"]), ("v", float(z["loss_v"]), case["gv"])):
        val, grads = eval_phase(s, case, phase)
        assert val.dtype == torch.float64
        assert abs(val.item() - gold_l) <= 1e-4 * abs(gold_l) + 1e-6
        comp = val.components
        assert abs(comp["I"].item() - float(z["I"])) <= 1e-4 * abs(float(z["I"]))
        for a, b in zip(grads, gold_g):
            assert a.dtype == np.float64 and G.rel(a, b) < 1e-3


def general_forms_agree_with_structure(s, case, device="cpu"):
    """per-path a / b and a callable c that HAPPEN to be constant / affine must give what the structured forms give
    (same kernels, other operand paths): pins the per-path b form, which no reference golden can (np.sum ambiguity of
    src/loss.py:69)"""
    z = case["z"]
    X, XV, BX = (torch.from_numpy(z[k]).to(device) for k in ("X", "XV", "BX"))
    d, n = s.setup["dim"], X.shape[0]
    g_ = torch.Generator().manual_seed(5)
    A = (torch.eye(d) + 0.3 * torch.randn(d, d, generator=g_)).float()
    B = (0.4 * torch.randn(d, generator=g_)).float()
    dom = s.new_domain()
    out = {}
    for form in ("structure", "general"):
        if form == "structure":
            a, b, c = xw.CoefA(A), xw.CoefB(B), xw.CoefC(0.25, -1.5)
        else:
            a = xw.CoefA(per_path=A.unsqueeze(0).expand(n, d, d).contiguous())
            b = xw.CoefB(per_path=B.unsqueeze(0).expand(n, d).contiguous())
            c = xw.CoefC(func=lambda X_, u: 0.25 - 1.5 * u)
        for phase in ("u", "v"):
            s.optimizer_u.zero_grad(); s.optimizer_v.zero_grad()
            pv, pu = s.v_net(XV), s.u_net(X)
            h, f, g = s.func_h(X[:, 0, :]), s.func_f(X), s.func_g(BX)
            L = xw.loss(s.config["alpha"], a, b, c, h, f, g, s.setup, dom, device)
            val = L.u(pu, pv, s.u_net, X, XV, BX) if phase == "u" else L.v(pu, pv, X, XV)
            val.backward()
            net = s.u_net if phase == "u" else s.v_net
            out[form, phase] = (val.item(), [q.grad.detach().cpu().numpy().copy() for q in net.parameters()])
    for phase in ("u", "v"):
        (l0, g0), (l1, g1) = out["structure", phase], out["general", phase]
        # (two evaluations on the GPU differ by fp32 rounding of v -- the order of the tensor-core accumulation varies from
        # run to run, xw_capi.cu tc_split_issue -- which the cancellation in I amplifies to a few 1e-6 of the loss)
        assert abs(l0 - l1) <= 5e-5 * abs(l0)
        for x, y in zip(g0, g1):
            assert G.rel(y, x) < 2e-4 or np.linalg.norm(x) == 0.0


def test_general_coefficient_forms_agree_with_structure(emu):
    case = G.load("cube_d4_ex43")
    s, _ = make_solver(case)
    general_forms_agree_with_structure(s, case)


def test_state_dict_names_match_reference_layout(emu):
    case = G.load("cube_d5_shipped_small")
    s, _ = make_solver(case)
    ku = list(s.u_net.state_dict().keys())
    assert ku[:6] == ["module.initial_layers.0.weight", "module.initial_layers.0.bias", "module.initial_layers.2.weight",
                      "module.initial_layers.2.bias", "module.initial_layers.4.weight", "module.initial_layers.4.bias"]
    assert "module.ODE_rhs.net.16.weight" in ku and "module.final_linear.bias" in ku
    assert [n for n, _ in s.u_net.named_parameters()][6:10] == [
        "module.ODE_rhs.net.0.weight", "module.ODE_rhs.net.0.bias", "module.ODE_rhs.net.2.weight", "module.ODE_rhs.net.2.bias"]
    kv = list(s.v_net.state_dict().keys())
    assert "module.hidden.weight" in kv and "module.net.2.weight" in kv and "module.net.20.weight" in kv
    assert [tuple(q.shape) for q in s.v_net.parameters()] == [(50, 6), (50,), (50, 50), (50,), (1, 50), (1,)]
    assert all(q.dtype == torch.float64 for q in s.u_net.parameters())


def test_hypercube_sampler_reproduces_reference_stream():
    """same seeds -> the golden samples of the unmodified reference, bit for bit"""
    case = G.load("cube_d5_shipped_small")
    p, z = case["params"], case["z"]
    torch.manual_seed(case["meta"]["seed"])
    np.random.seed(case["meta"]["seed"])
    # the reference solver constructor builds one throw-away domain and initialises both nets first
    s, _ = make_solver_for_rng(case)
    dom = s.new_domain()
    pts = xw.Comb_loader(p["N_r"], p["N_b"], dom, "cpu")
    X, XV, BX = pts[0]
    assert torch.equal(X, torch.from_numpy(z["X"]))
    assert torch.equal(XV, torch.from_numpy(z["XV"]))
    assert torch.equal(BX, torch.from_numpy(z["BX"]))
    # and the xavier initialisation consumed the RNG identically: weights equal the golden ones
    for q, w in zip(s.u_net.parameters(), case["thu_list"]):
        assert torch.equal(q.detach(), torch.from_numpy(np.asarray(w)))
    for q, w in zip(s.v_net.parameters(), case["thv_list"]):
        assert torch.equal(q.detach(), torch.from_numpy(np.asarray(w)))


def make_solver_for_rng(case):
    p = dict(case["params"])
    p["domain"] = case["meta"].get("domain_class", "Hypercube")
    prob = xw.problems.by_name(case["meta"]["funcs"], p["dim"])
    s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cpu",
                           "./", func_u_sol=prob.func_u_sol, p=2, log_json=False)
    return s, prob


def test_problem_callables_match_reference_values():
    case = G.load("cube_d4_ex43")
    z = case["z"]
    prob = xw.problems.ex4_3(4)
    X, BX = torch.from_numpy(z["X"]), torch.from_numpy(z["BX"])
    assert np.allclose(prob.func_h(X[:, 0, :]).numpy(), z["h"], rtol=2e-6, atol=1e-6)
    assert np.allclose(prob.func_f(X).numpy(), z["f"], rtol=2e-5, atol=2e-5)
    assert np.allclose(prob.func_g(BX).numpy(), z["g"], rtol=2e-6, atol=1e-6)


def test_unsupported_inputs_raise(emu):
    case = G.load("cube_d3_small_nets")
    s, prob = make_solver(case)
    z = case["z"]
    X = torch.from_numpy(z["X"])
    # (coefficients that vary over the sample or are not affine in u are no longer refused: they classify as
    # per-path / callable forms, test_coefficient_classification)
    with pytest.raises(RuntimeError):
        xw.NeuralODE(20, 1, prob.func_h, prob.func_g, s.setup, 10, 8, s.new_domain(), solver="dopri5")
    with pytest.raises(RuntimeError):
        xw.NeuralODE(20, 1, prob.func_h, prob.func_g, s.setup, 10, 8, s.new_domain(), adjoint=True)
    with pytest.raises(TypeError):
        xw.loss(1.0, torch.zeros(3, 3, 4, 7), xw.CoefB(), xw.CoefC(), None, None, None, s.setup, s.new_domain(), "cpu")


def test_product_path_refuses_cpu_tensors():
    """without the test injection the product path must fail loudly on CPU tensors / missing GPU"""
    case = G.load("cube_d3_small_nets")
    s, _ = make_solver(case)
    with pytest.raises(RuntimeError):
        eval_phase(s, case, "u")


def test_collapsed_layout_equals_dense_layout(emu):
    """CollapsedPaths(times, x) through the public API == the reference's repeated [N, L, C] tensors"""
    case = G.load("cube_d5_alpha1_randbias")
    s, _ = make_solver(case)
    z = case["z"]
    dense = {k: torch.from_numpy(z[k]) for k in ("X", "XV", "BX")}
    col = {k: xw.CollapsedPaths(v[0, :, 0].clone(), v[:, 0, 1:].clone()) for k, v in dense.items()}
    assert torch.equal(col["X"].dense(), dense["X"])
    assert torch.equal(col["X"][:, :, 2], dense["X"][:, :, 2]) and torch.equal(col["X"][:, 0, :], dense["X"][:, 0, :])
    outs = []
    for data in (dense, col):
        res = []
        for phase in ("u", "v"):
            dom = s.new_domain()
            s.optimizer_u.zero_grad(); s.optimizer_v.zero_grad()
            pv, pu = s.v_net(data["XV"]), s.u_net(data["X"])
            h, f, g, a, b, c = xw.func_eval(data["X"], data["BX"], s.setup, pu, s.func_a, s.func_b, s.func_c, s.func_h,
                                            s.func_f, s.func_g)
            L = xw.loss(s.config["alpha"], a, b, c, h, f, g, s.setup, dom, "cpu")
            val = L.u(pu, pv, s.u_net, data["X"], data["XV"], data["BX"]) if phase == "u" else L.v(pu, pv, data["X"], data["XV"])
            val.backward()
            net = s.u_net if phase == "u" else s.v_net
            res.append((val.item(), [q.grad.clone() for q in net.parameters()]))
        outs.append(res)
    for (l0, g0), (l1, g1) in zip(*outs):
        assert abs(l0 - l1) <= 1e-9 * abs(l0)
        for a_, b_ in zip(g0, g1):      # cross-warp fp32 atomics in the tiled backward: order-dependent rounding
            assert G.rel(a_.numpy(), b_.numpy()) < 1e-5
    # forward-only evaluation agrees too
    with torch.no_grad():
        assert torch.allclose(s.u_net(dense["X"]), s.u_net(col["X"]), atol=1e-6)


def test_train_iteration_runs_and_updates_both_nets(emu):
    case = G.load("cube_d3_small_nets")
    s, _ = make_solver(case)
    torch.manual_seed(0)
    dom = s.new_domain()
    pts = xw.Comb_loader(16, 12, dom, "cpu")
    before = [q.detach().clone() for q in list(s.u_net.parameters()) + list(s.v_net.parameters())]
    lu, lv = s.train_iteration(dom, pts)
    assert torch.isfinite(lu) and torch.isfinite(lv)
    after = list(s.u_net.parameters()) + list(s.v_net.parameters())
    assert all(not torch.equal(a, b) for a, b in zip(before, after))


def test_test_function_cache_gives_the_same_training_step(emu):
    """2nd u-step and the v-step of an iteration reuse the cached v-net values (same sample, same
    theta_v): the parameters after one full iteration must equal those of the uncached path"""
    case = G.load("cube_d3_small_nets")
    outs = []
    for reuse in (True, False):
        s, _ = make_solver(case)
        s.reuse_v = reuse
        torch.manual_seed(0)
        dom = s.new_domain()
        pts = xw.Comb_loader(24, 16, dom, "cpu")
        calls0 = dict(xw.hotpath.CALLS)
        s.train_iteration(dom, pts)
        s.train_iteration(dom, xw.Comb_loader(24, 16, dom, "cpu"))
        outs.append([q.detach().clone() for q in list(s.u_net.parameters()) + list(s.v_net.parameters())])
    for a, b in zip(*outs):
        assert G.rel(a.numpy(), b.numpy()) < 1e-5


def test_single_time_group_matches_reference_golden(emu):
    """first interior group of the sphere domains (one time point at T0): the reference's rank-2
    shortcut + [n, n] broadcasts, reproduced on the host side"""
    case = G.load("cone_d5_g0_single_time")
    s, _ = make_solver(case)
    z = case["z"]
    for phase, gold_l, gold_g in (("u", float(z["loss_u"]), case["gu"]), ("v", float(z["loss_v"]), case["gv"])):
        val, grads = eval_phase(s, case, phase)
        assert abs(val.item() - gold_l) <= 1e-6 * abs(gold_l)
        assert abs(val.components["I"].item() - float(z["I"])) <= 1e-6 * abs(float(z["I"]))
        for a, b in zip(grads, gold_g):
            assert G.rel(a, b) < 1e-6 or np.linalg.norm(b) == 0.0


def test_sphere_samplers_reproduce_reference_groups():
    """same seeds -> the group shapes / first times the unmodified reference produced (SURVEY Appendix B
    protocol; the bit-for-bit comparison against the reference's samplers was done in the build container)"""
    for name, golden in (("NSphere_TCone", "cone_d5_g2"), ("NSphere_THourglass", "hourglass_d5_g4_reentry")):
        case = G.load(golden)
        seed = case["meta"]["seed"]
        torch.manual_seed(seed)
        np.random.seed(seed)
        s, _ = make_solver_for_rng(case)       # consumes the RNG exactly like the reference constructor
        dom = s.new_domain()
        pts = xw.Comb_loader(300, 300, dom, "cpu")
        X, XV, BX = pts[case["meta"]["group"]]
        z = case["z"]
        # spatial coordinates and grid times bit for bit; the hourglass' computed entry time |x| / r (row 0 of a
        # re-entry segment) to the last bit or two: torch's vectorised sum over the d coordinates associates
        # differently on AVX2 and AVX-512 hosts, and the golden comes from whichever host generated it
        for mine, ref in ((X, z["X"]), (XV, z["XV"]), (BX, z["BX"])):
            ref = torch.from_numpy(ref)
            assert mine.shape == ref.shape and torch.equal(mine[..., 1:], ref[..., 1:])
            assert torch.equal(mine[:, 1:, 0], ref[:, 1:, 0])
            assert float((mine[:, 0, 0] - ref[:, 0, 0]).abs().max()) <= 4.5e-16


def test_training_iteration_on_cone_domain(emu):
    """one outer iteration over every variable-length group of NSphere_TCone, including the
    single-time-point group and the single-time boundary groups"""
    case = G.load("cone_d5_g2")
    s, _ = make_solver(case)
    s.setup["N_r"], s.setup["N_b"] = 60, 60
    torch.manual_seed(1)
    np.random.seed(1)
    dom = s.new_domain()
    pts = xw.Comb_loader(60, 60, dom, "cpu")
    assert len(pts) > 2
    before = [q.detach().clone() for q in list(s.u_net.parameters()) + list(s.v_net.parameters())]
    lu, lv = s.train_iteration(dom, pts)
    assert torch.isfinite(lu) and torch.isfinite(lv)
    after = list(s.u_net.parameters()) + list(s.v_net.parameters())
    assert all(torch.isfinite(a).all() for a in after)
    assert any(not torch.equal(a, b) for a, b in zip(before, after))


def test_product_library_carries_tcgen05_code():
    """the shipped .so holds tensor-core (tcgen05) SASS for the test-function kernels: UTCHMMA = tcgen05.mma,
    LDTM / STTM = tcgen05.ld / st (B200_PROFILING.md).  Checked with cuobjdump on the build box (no GPU needed)."""
    import shutil
    import subprocess
    from xnode_wan_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe) or not os.path.exists(_lib.LIB_PATH):
        pytest.skip("cuobjdump or the product library is not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "STTM"):
        assert mnemonic in sass, mnemonic


PAD_CASES = ["pad_cube_d3_rk4", "pad_cube_d5_midpoint", "pad_cone_d5", "pad_hourglass_d5_early", "pad_hourglass_d5_late",
             "pad_hourglass_d5_late_onbdry"]


@pytest.mark.parametrize("name", PAD_CASES[3:])
def test_hourglass_bound_pad_builds_the_reference_groups(name):
    """per-path entry times, filled grids, grouping by grid length and the first-path-represents-the-group rule of the
    reference's hourglass bound_pad (src/dataset.py:127-152); groups / positions / grids stored by the unmodified
    reference (tests/golden/make_pad_golden.py)"""
    z = np.load(os.path.join(G.GOLDEN_DIR, "extra", name + ".npz"))
    case = G.load(str(z["base"]))
    s, _ = make_solver_for_rng(case)
    dom = s.new_domain()
    dom.times = torch.from_numpy(z["dom_times"])
    path_i, pos, grids = dom.bound_pad(torch.from_numpy(z["X"]))
    assert len(grids) == int(z["n_groups"])
    assert np.array_equal(np.concatenate([q.numpy() for q in path_i]), z["order"])
    for k in range(len(grids)):
        assert np.array_equal(pos[k].numpy(), z["pos%d" % k])
        assert np.allclose(grids[k].numpy(), z["grid%d" % k], rtol=0, atol=1e-15)


@pytest.mark.parametrize("name", PAD_CASES)
def test_evaluation_from_inside_the_domain_matches_reference(emu, name):
    """u_net(X) for paths that start inside the domain after T0: the reference pads the time grid back to T0 and
    fills the gaps (bound_pad / fillt, src/model.py:92-94, src/dataset.py:13-32); golden vectors from the unmodified
    reference (tests/golden/make_pad_golden.py)"""
    z = np.load(os.path.join(G.GOLDEN_DIR, "extra", name + ".npz"))
    case = G.load(str(z["base"]))
    s, _ = make_solver(case)
    with torch.no_grad():
        u = s.u_net(torch.from_numpy(z["X"]))
    assert u.shape == z["u"].shape + (1,)       # (the early-time hourglass batch also returns u at T0: reference quirk)
    assert np.abs(u.numpy()[..., 0] - z["u"]).max() < 2e-5


def test_fillt_grid_properties():
    from xnode_wan_b200.dataset import fillt
    t = torch.tensor([0.0, 0.3, 0.35, 0.9])
    pos, grid = fillt(t, 1.0, 0.0, 20)
    assert torch.allclose(grid[pos], t)                        # the requested times sit where `pos` says
    assert bool((grid[1:] > grid[:-1]).all())                  # strictly increasing
    assert float((grid[1:] - grid[:-1]).max()) <= 0.05 + 1e-6  # no step larger than (T - T0) / min_steps
    pos1, grid1 = fillt(torch.tensor([0.0, 0.01, 0.02]), 1.0, 0.0, 20)
    assert grid1.numel() == 1                                  # the reference's degenerate case is reproduced


def test_coefficient_cache_is_keyed_on_the_callables(emu):
    """two problems whose closures could share an id() after garbage collection must not share the
    classified structure; the probe looks at a spread of paths, not the first four"""
    import gc
    case = G.load("cube_d3_small_nets")
    s, prob = make_solver(case)
    X = torch.from_numpy(case["z"]["X"])
    seen = []
    for k in range(4):
        A = torch.eye(3) * (k + 1.0)
        fa = (lambda A_: (lambda X_, i, j: torch.full(X_.shape[:-1], float(A_[i, j]))))(A)
        a, b, c = xw.training.classify_coefficients(X, s.setup, fa, prob.func_b, prob.func_c)
        seen.append(None if a[0] == "identity" else float(a[1][0, 0]))
        del fa
        gc.collect()
    assert seen == [None, 2.0, 3.0, 4.0]
    # a coefficient that differs only on the LAST path is caught (the old probe read paths 0..3 only)
    last = X.shape[0] - 1
    assert last >= 4

    def fa_var(X_, i, j):
        out = torch.ones(X_.shape[:-1]) if i == j else torch.zeros(X_.shape[:-1])
        if X_.shape[0] > 0 and i == j:
            out = out + (X_[:, :, 1] == X[last, 0, 1]).float()
        return out
    assert xw.training.classify_coefficients(X, s.setup, fa_var, prob.func_b, prob.func_c)[0] == ("per_path",)


def test_coefficient_classification(emu):
    """constant / affine callables become structure, everything else the per-path / callable forms; func_eval then
    evaluates a, b on time-row 0 of every path (reference: dense a[d,d,N,L], src/training.py:32-41)"""
    from tests import _general_coef as GC
    case = G.load("cube_d3_small_nets")
    s, prob = make_solver(case)
    z = case["z"]
    X, BX = torch.from_numpy(z["X"]), torch.from_numpy(z["BX"])
    kinds = xw.training.classify_coefficients(X, s.setup, prob.func_a, prob.func_b, prob.func_c)
    assert kinds[0] == ("identity",) and kinds[1] == ("zero",) and kinds[2][0] == "affine"
    assert xw.training.classify_coefficients(X, s.setup, prob.func_a, prob.func_b, lambda X_, u: -u * u)[2] == ("callable",)
    assert xw.training.classify_coefficients(X, s.setup, prob.func_a, prob.func_b, lambda X_, u: X_[..., 1:2] + 0 * u)[2] == ("callable",)
    fb = lambda X_, i: 0.5 * X_[..., 1 + i]
    h, f, g, a, b, c = xw.func_eval(X, BX, s.setup, None, GC.func_a, fb, GC.func_c, prob.func_h, prob.func_f, prob.func_g)
    d = s.setup["dim"]
    assert a.matrix is None and tuple(a.per_path.shape) == (X.shape[0], d, d)
    for i in range(d):
        for j in range(d):
            assert torch.allclose(a.per_path[:, i, j], GC.func_a(X, i, j)[:, 0].float())
        assert torch.allclose(b.per_path[:, i], fb(X, i)[:, 0].float())
    assert c.func is GC.func_c


def test_vcache_grows_with_the_batch_and_capi_checks_capacity(emu):
    """the test-function cache is sized per batch; a later, larger batch re-allocates it (and the C ABI
    refuses a buffer that is too small instead of writing past it)"""
    import ctypes as C
    case = G.load("cube_d3_small_nets")
    s, prob = make_solver(case)
    torch.manual_seed(3)
    np.random.seed(3)
    dom = s.new_domain()
    small = xw.Comb_loader(16, 12, dom, "cpu")
    big = xw.Comb_loader(40, 12, dom, "cpu")
    s.sub_step("u", dom, small)
    n_small = s._vc_buf.numel()
    lu = s.sub_step("u", dom, big)
    assert s._vc_buf.numel() > n_small
    # same numbers as a solver that only ever saw the big batch
    s2, _ = make_solver(case)
    with torch.no_grad():
        pass
    s2.sub_step("u", dom, small)        # same optimiser history
    s3, _ = make_solver(case)
    s3.sub_step("u", dom, small)
    lu3 = s3.sub_step("u", dom, big)
    assert lu.item() == lu3.item()
    # raw C ABI: capacity one float short -> error, nothing written
    from tests import _lowlevel as LL
    dims = LL.make_dims(case)
    z = case["z"]
    N, L, Cc = z["X"].shape
    need = emu.cdll.xw_vcache_floats(C.byref(dims), N, L)
    be = LL.NumpyBackend()
    Xd, XVd = be.arr(z["X"]), be.arr(z["XV"])
    thu, thv = be.arr(LL.flat_theta(case["thu_list"])), be.arr(LL.flat_theta(case["thv_list"]))
    pts = xw._lib.Points(be.ptr(XVd).value, L * Cc, Cc, be.ptr_off(XVd, 1).value, L * Cc, Cc)
    dom_c = LL.make_domain(case["meta"]["domain"])
    coef = xw._lib.Coef(0.0, -1.0, None, None)
    wsb = emu.workspace_bytes(dims, N, L)
    ws, sums = be.zeros(wsb, np.uint8), be.zeros(8, np.float64)
    cu, cv, vc = be.zeros(N * L), be.zeros(N * L), be.zeros(need)
    with pytest.raises(xw._lib.XwError, match="cache too small"):
        emu.call("xw_interior_forward", C.byref(dims), C.byref(dom_c), C.byref(coef), be.ptr(thu), be.ptr(thv),
                 be.ptr_off(Xd, 1), L * Cc, be.ptr(be.arr(z["X"][0, :, 0])), L, C.byref(pts), be.ptr(be.arr(z["h"])),
                 be.ptr(be.arr(z["grad_h"])), be.ptr(be.arr(z["f"])), N, be.ptr(sums), be.ptr(cu), be.ptr(cv), None,
                 be.ptr(ws), wsb, None, None, be.ptr(vc), 1, None, need - 1, 0)
    assert not sums.any() and not vc.any()


def test_coefficient_values_are_evaluated_once_per_sample(emu):
    """the sub-steps of one outer iteration run on one sample: h, f, g, grad h are evaluated by the first one only, and
    the numbers are the same as re-evaluating them every time"""
    case = G.load("cube_d3_small_nets")
    calls = {"h": 0, "f": 0}

    def run(cache):
        torch.manual_seed(4)
        np.random.seed(4)
        s, prob = make_solver(case)
        fh, ff = s.func_h, s.func_f
        s.func_h = lambda X0: (calls.__setitem__("h", calls["h"] + 1), fh(X0))[1]
        s.func_f = lambda X: (calls.__setitem__("f", calls["f"] + 1), ff(X))[1]
        s.u_net.module.h = s.func_h
        dom = s.new_domain()
        pts = xw.Comb_loader(32, 24, dom, "cpu")
        out = []
        if not cache:
            orig = s._step
            s._step = lambda ph, d_, b_, vp=(None, 0), token=None: orig(ph, d_, b_, vp, token=None)
        for _ in range(2):
            lu, lv = s.train_iteration(dom, pts)
            out.append((lu.item(), lv.item()))
        return out
    calls.update(h=0, f=0)
    a = run(True)
    n_cached = dict(calls)
    calls.update(h=0, f=0)
    b = run(False)
    for (x0, y0), (x1, y1) in zip(a, b):       # (the kernels' shared-memory atomics make runs agree to ~1e-8, not bit for bit)
        assert abs(x0 - x1) <= 1e-6 * abs(x0) and abs(y0 - y1) <= 1e-6 * abs(y0)
    assert n_cached["f"] == 1 and calls["f"] == 6          # 2 iterations x 3 sub-steps on ONE sample
    assert n_cached["h"] < calls["h"]


def test_adam_step_kernel_matches_torch_adam(emu):
    """xw_adam_step on a flat fp64 vector == torch.optim.Adam (defaults) on the same numbers, step after step"""
    import ctypes as C
    torch.manual_seed(0)
    n = 1651
    p0 = torch.randn(n, dtype=torch.float64)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=0.015)
    p, m, v = p0.clone(), torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)
    step = torch.zeros(1, dtype=torch.int64)
    p32 = torch.zeros(n, dtype=torch.float32)
    ptr = lambda t: C.c_void_p(t.data_ptr())   # noqa: E731
    for it in range(6):
        g = (torch.randn(n) * (10.0 ** (it - 3))).float()
        ref.grad = g.double()
        opt.step()
        emu.call("xw_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), ptr(step), ptr(p32), n, 0.015, 0.9, 0.999, 1e-8, None)
        assert int(step) == it + 1
        assert torch.allclose(p, ref.detach(), rtol=1e-13, atol=1e-15)
        assert torch.equal(p32, p.float())


def test_fused_optimizer_follows_torch_adam(emu):
    """NODE_WAN_solver(fused_optimizer=True): flat parameter buffers + single-launch Adam; same trajectory as the two
    torch.optim.Adam of the reference, and state_dict / named_parameters keep the reference's names and shapes"""
    case = G.load("cube_d3_small_nets")

    def run(fused):
        torch.manual_seed(9)
        np.random.seed(9)
        p = dict(case["params"])
        p["domain"] = "Hypercube"
        prob = xw.problems.ex4_1()
        s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cpu", "./",
                               func_u_sol=prob.func_u_sol, p=2, log_json=False, fused_optimizer=fused)
        names = [(k, tuple(v.shape)) for k, v in s.u_net.state_dict().items()]
        dom = s.new_domain()
        out = []
        for _ in range(3):
            pts = xw.Comb_loader(p["N_r"], p["N_b"], dom, "cpu")
            lu, lv = s.train_iteration(dom, pts)
            out.append((lu.item(), lv.item()))
        return out, names, [q.detach().clone() for q in list(s.u_net.parameters()) + list(s.v_net.parameters())], s
    a, na, pa, sa = run(False)
    b, nb, pb, sb = run(True)
    assert sb.fused_optimizer and not sa.fused_optimizer and na == nb
    for (x0, y0), (x1, y1) in zip(a, b):
        assert abs(x0 - x1) <= 1e-6 * abs(x0) and abs(y0 - y1) <= 1e-6 * max(abs(y0), 1.0)
    for x, y in zip(pa, pb):
        assert torch.allclose(x, y, rtol=1e-6, atol=1e-8)
    # parameters modified in place behind the optimiser's back are picked up (version counters)
    with torch.no_grad():
        for q in sb.u_net.parameters():
            q.mul_(0.5)
    th = xw.hotpath.flatten_params(sb.u_net.module.kernel_parameters())
    assert torch.equal(th, torch.cat([q.detach().reshape(-1) for q in sb.u_net.module.kernel_parameters()]).float())


def test_on_device_cube_sampler_distribution():
    """`sample_device` draws with torch's generator of that device (same distributions as the reference sampler,
    src/dataset.py:251-276, another RNG stream): interior uniform in the cube, boundary = int(N_b/d/2) paths per face with
    exactly one coordinate pinned to top / bot (the remainder goes to the last face), random order, shared time grid with
    pinned end points; collapsed layout == dense layout"""
    torch.manual_seed(0)
    d, N, Nb = 6, 20000, 1203
    dom = xw.Hypercube((-1.0, 2.0), d, 0.0, 1.0, 20, sample_device="cpu", collapsed=True)
    X, B = dom.interior(N), dom.boundary(Nb)
    assert isinstance(X, xw.CollapsedPaths) and X.shape == (N, 20, d + 1) and B.shape == (Nb, 20, d + 1)
    t = dom.times
    assert float(t[0]) == 0.0 and float(t[-1]) == 1.0 and bool(torch.all(t[1:] >= t[:-1]))
    x = X.x
    assert float(x.min()) >= -1.0 and float(x.max()) < 2.0
    assert abs(float(x.mean()) - 0.5) < 0.02 and abs(float(x.var()) - 9.0 / 12.0) < 0.02       # U(-1, 2)
    hist = torch.histc(x[:, 0], bins=10, min=-1.0, max=2.0) / N
    assert float((hist - 0.1).abs().max()) < 0.012
    xb = B.x
    on_top, on_bot = (xb == 2.0), (xb == -1.0)
    pinned = on_top | on_bot
    assert bool(torch.all(pinned.sum(1) == 1))                     # exactly one pinned coordinate per boundary path
    per_face = Nb // d // 2
    top_counts, bot_counts = on_top.sum(0).tolist(), on_bot.sum(0).tolist()
    assert top_counts == [per_face] * d
    assert bot_counts[:-1] == [per_face] * (d - 1) and bot_counts[-1] == Nb - per_face * (2 * d - 1)
    free = xb[~pinned]
    assert float(free.min()) > -1.0 and float(free.max()) < 2.0 and abs(float(free.mean()) - 0.5) < 0.05
    # the permutation mixes the faces: the first block of paths is not all on face 0
    assert int(on_top[:per_face, 0].sum()) < per_face
    dense = X.dense()
    assert torch.equal(dense[:, 3, 1:], x) and torch.equal(dense[5, :, 0], t)


@pytest.mark.parametrize("cls", ["NSphere_TCone", "NSphere_THourglass"])
def test_on_device_sphere_samplers_build_the_reference_groups(cls):
    """the vectorised group construction used by `sample_device` (torch ops on the sampling device) yields, for the SAME
    points, exactly the groups of the reference-order CPU sampler (src/dataset.py:81-117, :185-214); and its own draws
    have the right structure (inside the ball, boundary counts int(N_b scale(t)^d))"""
    Dom = getattr(xw, cls)
    torch.manual_seed(2)
    np.random.seed(2)
    dom = Dom(1.0, 5, 0.0, 1.0, 20)
    st = np.random.get_state()
    pts = dom._ball(400)                                   # [dim, N] as the CPU sampler draws them
    np.random.set_state(st)
    ref = dom.interior(400)
    got = dom._groups_from(torch.from_numpy(pts).transpose(0, 1).contiguous())
    assert [tuple(g.shape) for g in got] == [tuple(g.shape) for g in ref]
    for a, b in zip(got, ref):
        assert a.dtype == b.dtype == torch.float64 and torch.equal(a, b)
    dev = Dom(1.0, 5, 0.0, 1.0, 20, times=dom.times, sample_device="cpu")
    groups = dev.interior(3000)
    assert sum(g.shape[0] for g in groups if float(g[0, 0, 0]) == 0.0) == 3000        # every path has a segment from T0
    lens = [g.shape[1] for g in groups]
    assert lens == sorted(lens)
    for g in groups:
        assert bool(torch.all(dev.func_w(g) > 0) if cls == "NSphere_TCone" else torch.all(dev.func_w(g[:, 1:] if float(g[0, 0, 0]) != 0.0 else g) > 0))
    bd = dev.boundary(500)
    want = [int(500 * dev._radius_scale(t) ** 5) for t in dev.times.numpy()]
    assert [g.shape[0] for g in bd] == [n for n in want if n]
    for g in bd:
        assert g.shape[1] == 1 and float(dev.func_w(g).abs().max()) < 1e-9           # on the sphere of that time


def test_error_norms_keep_the_solution_values_with_the_sample(emu):
    """stop() -> rel_err runs after every u sub-iteration on the same sample (reference src/training.py:142): func_u_sol
    is evaluated once per sample, the numbers equal the straightforward evaluation, an in-place edit of the sample drops
    the kept values"""
    case = G.load("cube_d3_small_nets")
    s, prob = make_solver(case)
    torch.manual_seed(2)
    dom = s.new_domain()
    X = dom.interior(40)
    calls = [0]
    sol = lambda Z: (calls.__setitem__(0, calls[0] + 1), prob.func_u_sol(Z))[1]

    def plain(Z):
        pred = s.u_net(Z).materialize().squeeze()
        u = prob.func_u_sol(Z)
        return ((dom.V() * torch.mean(torch.abs(u - pred) ** 2)) ** 0.5) / ((dom.V() * torch.mean(torch.abs(u) ** 2)) ** 0.5)
    r1 = xw.rel_err(X, s.u_net, sol, 2, dom.V(), 40)
    r2 = xw.rel_err(X, s.u_net, sol, 2, dom.V(), 40)
    assert calls[0] == 1
    assert float(r1) == float(r2) == float(plain(X))
    X[:, :, 1].mul_(0.5)                    # the sample changes in place: version counter moves
    r3 = xw.rel_err(X, s.u_net, sol, 2, dom.V(), 40)
    assert calls[0] == 2 and float(r3) == float(plain(X)) and float(r3) != float(r1)
    other = lambda Z: prob.func_u_sol(Z) * 2.0          # another callable on the same sample is not served from the cache
    n = xw.L_norm(X, s.u_net, 2, other, dom.V(), 40, error=False)
    assert abs(float(n) - 2.0 * float(xw.L_norm(X, s.u_net, 2, sol, dom.V(), 40, error=False))) < 1e-12


def test_train_draws_its_samples_in_the_reference_order(emu, monkeypatch):
    """train() draws the next iteration's domain and sample while the GPU still runs the v-steps; the ORDER of the draws
    (domain_k, sample_k, logging sample_k, domain_k+1, ...; reference src/training.py:114-115, 165) must not change, or a
    seed would no longer give the reference's samples"""
    case = G.load("cube_d3_small_nets")
    seen = []

    def record_train(log_l2):
        s, _ = make_solver(case)
        s.iterations, s.keep_l2_history = 3, log_l2
        torch.manual_seed(11)
        seen.clear()
        real_domain = s.new_domain
        s.new_domain = lambda **kw: (seen.append(("domain",)), real_domain(**kw))[1]
        real_loader = xw.training.Comb_loader

        class Loader(real_loader):
            def __init__(self, *a):
                super().__init__(*a)
                seen.append(("sample", float(self.interioru[0, 0, 1]), float(self.boundary[-1, 0, 2])))
        monkeypatch.setattr(xw.training, "Comb_loader", Loader)
        s.train()
        monkeypatch.setattr(xw.training, "Comb_loader", real_loader)
        return list(seen)

    def sequential(log_l2):
        s, _ = make_solver(case)
        torch.manual_seed(11)
        out = []
        n_r, n_b = s.local_counts()
        for k in range(3):
            dom = s.new_domain()
            out.append(("domain",))
            for _ in range(2 if log_l2 else 1):
                p = xw.Comb_loader(n_r, n_b, dom, "cpu")
                out.append(("sample", float(p.interioru[0, 0, 1]), float(p.boundary[-1, 0, 2])))
        return out
    for log_l2 in (True, False):
        assert record_train(log_l2) == sequential(log_l2)


def test_loss_scalars_entry_equals_the_torch_arithmetic(emu):
    """xw_loss_scalars (one launch) against hotpath.loss_from_sums + the coefficient formulas it replaced"""
    import ctypes as C
    from types import SimpleNamespace
    hp = xw.hotpath
    g = torch.Generator().manual_seed(5)
    for phase, nb in (("u", 37), ("v", 0), ("u", 0)):
        sums = torch.rand(8, dtype=torch.float64, generator=g) + 0.1
        sums[0] -= 0.6
        V, N, L, Lb, alpha, side = 32.0, 4000, 20, 20, 1e8, 1.0
        out = torch.empty(8, dtype=torch.float64)
        rc = emu.cdll.xw_loss_scalars(C.c_void_p(sums.data_ptr()), 0 if phase == "u" else 1, V, float(N), float(L), float(nb),
                                      float(Lb), alpha, side, C.c_void_p(out.data_ptr()), None)
        assert rc == 0
        b = SimpleNamespace(N_glob=N, L=L, Nb_glob=nb, Lb=Lb)
        I, S, init, bdry, integ = hp.loss_from_sums(sums, b, V, alpha)
        want = [integ + alpha * (init + bdry) if phase == "u" else -integ, I, S, init, bdry,
                (2.0 / I) * (V / (N * L)) * (1 if phase == "u" else -1), torch.tensor(2.0 * alpha / N) if phase == "u" else 2.0 / sums[3],
                torch.tensor(side)]
        for a, w in zip(out.tolist(), want):
            assert abs(a - float(w)) <= 1e-14 * abs(float(w)) + 1e-300
    bad = emu.cdll.xw_loss_scalars(C.c_void_p(sums.data_ptr()), 2, 1.0, 1.0, 1.0, 0.0, 1.0, 1.0, 1.0, C.c_void_p(out.data_ptr()), None)
    assert bad != 0
