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
    xw.hotpath._TEST_ALLOW_HOST = True
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
    assert lib.cdll.xw_abi_version() == 1
