"""CPU: the N>1 path (paths sharded over ranks, all-reduced sums and gradients) with world_size 2 on
the gloo backend and the emulated kernels: every rank must obtain the single-process loss and
gradients."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from tests import _golden as G


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    import xnode_wan_b200 as xw
    from tests.host_emu import build_emu
    from tests.test_host_api_emu import make_solver
    xw._lib._LIB = xw._lib.XwLib(build_emu.build())
    case = G.load(name)
    s, _ = make_solver(case)
    assert s.world == world and s.rank == rank
    z = case["z"]
    N, Nb = z["X"].shape[0], z["BX"].shape[0]
    sl = slice(rank * N // world, (rank + 1) * N // world)
    slb = slice(rank * Nb // world, (rank + 1) * Nb // world)
    X, XV, BX = torch.from_numpy(z["X"][sl]), torch.from_numpy(z["XV"][sl]), torch.from_numpy(z["BX"][slb])
    dom = s.new_domain()
    res = {}
    for phase in ("u", "v"):
        s.optimizer_u.zero_grad(); s.optimizer_v.zero_grad()
        pv, pu = s.v_net(XV), s.u_net(X)
        h, f, g, a, b, c = xw.func_eval(X, BX, s.setup, pu, s.func_a, s.func_b, s.func_c, s.func_h, s.func_f, s.func_g)
        L = xw.loss(s.config["alpha"], a, b, c, h, f, g, s.setup, dom, "cpu")
        L.N_glob, L.Nb_glob = N, Nb
        val = L.u(pu, pv, s.u_net, X, XV, BX) if phase == "u" else L.v(pu, pv, X, XV)
        val.backward()
        net = s.u_net if phase == "u" else s.v_net
        res[phase] = (val.item(), [q.grad.numpy().copy() for q in net.parameters()])
    out[rank] = res
    torch.distributed.destroy_process_group()


@pytest.mark.parametrize("name", ["cube_d5_alpha1_randbias", "cube_d4_ex43"])
def test_two_ranks_reproduce_single_process(name):
    case = G.load(name)
    z = case["z"]
    assert z["X"].shape[0] % 2 == 0 and z["BX"].shape[0] % 2 == 0
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), name, out), nprocs=2, join=True)
    for rank in (0, 1):
        for phase, gold_l, gold_g in (("u", float(z["loss_u"]), case["gu"]), ("v", float(z["loss_v"]), case["gv"])):
            val, grads = out[rank][phase]
            assert abs(val - gold_l) <= 1e-4 * abs(gold_l) + 1e-6
            for a, b in zip(grads, gold_g):
                assert G.rel(a, b) < 1e-3
    # both ranks hold the same numbers (replicas stay identical)
    for phase in ("u", "v"):
        assert out[0][phase][0] == out[1][phase][0]
        for a, b in zip(out[0][phase][1], out[1][phase][1]):
            assert np.array_equal(a, b)


def test_capi_exports_every_declared_symbol():
    """the product library loads on a CPU-only box and exports every function include/*.h declares"""
    import re
    import xnode_wan_b200 as xw
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "xnode_wan_b200.h")).read()
    declared = set(re.findall(r"\b(xw_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(xw._lib.EXPORTS), declared ^ set(xw._lib.EXPORTS)
    if not os.path.exists(xw._lib.LIB_PATH):
        import importlib
        importlib.import_module("xnode-wan-pde-solver_b200.build").build()
    lib = xw._lib.XwLib()
    for name in declared:
        assert hasattr(lib.cdll, name)
    assert lib.cdll.xw_abi_version() == xw._lib.ABI_VERSION


def _train_worker(rank, world, port, out):
    """NODE_WAN_solver.train() on 2 ranks with DIFFERENT caller seeds and a stop criterion that fires on rank 0 only"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    import xnode_wan_b200 as xw
    from tests.host_emu import build_emu
    xw._lib._LIB = xw._lib.XwLib(build_emu.build())
    torch.manual_seed(100 + rank)            # replicas would start from different weights without the broadcast
    np.random.seed(5)                        # ... and draw the same numpy stream without the per-rank reseed
    prob = xw.problems.ex4_1()
    p = xw.problems.cube_params(dim=3, N_r=64, N_b=48, N_t=5, iterations=3, alpha=10)
    calls = []

    def stop(s, points, domain):
        calls.append(1)
        return s.rank == 0 and len(calls) >= 3        # rank 1 never wants to stop
    cwd = os.getcwd()
    import tempfile
    os.chdir(tempfile.mkdtemp())
    try:
        s = xw.NODE_WAN_solver(p, prob.func_a, prob.func_b, prob.func_c, prob.func_h, prob.func_f, prob.func_g, "cpu",
                               "./", stop=stop, func_u_sol=prob.func_u_sol, p=2, log_json=True)
        w0 = [q.detach().clone() for q in s.u_net.parameters()]
        dom = s.new_domain()
        first = xw.Comb_loader(*s.local_counts(), dom, "cpu").interioru[:, 0, 1:].clone()
        hist = s.train(report=False)
        files = sorted(os.listdir("."))
    finally:
        os.chdir(cwd)
    out[rank] = dict(w0=[w.numpy() for w in w0], w1=[q.detach().numpy().copy() for q in s.u_net.parameters()],
                     first=first.numpy(), stopped=hist.get("stopped_at_subiter"), ncalls=len(calls), files=files)
    torch.distributed.destroy_process_group()


def test_two_rank_train_loop_stays_consistent():
    """train() on several ranks -- identical initial weights (broadcast), different shards (per-rank
    sampling streams), ONE stop decision (all-reduced), files written by rank 0 only"""
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    for x, y in zip(a["w0"], b["w0"]):
        assert np.array_equal(x, y)              # started identical although the callers seeded differently
    for x, y in zip(a["w1"], b["w1"]):
        assert np.array_equal(x, y)              # and stayed identical through the Adam steps
    assert not np.array_equal(a["first"], b["first"])     # the shards are different samples
    assert a["stopped"] == 3 and b["stopped"] == 3 and a["ncalls"] == b["ncalls"] == 3
    assert any(f.startswith("losses_NODE") for f in a["files"]) and "best_model_weights_NODE.pth" in a["files"]
    assert b["files"] == []                      # only rank 0 writes logs / checkpoints
