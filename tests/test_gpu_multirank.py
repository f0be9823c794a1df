"""Time-sharded EM path against the single-GPU run.  With >= `world` CUDA devices the ranks use one GPU each over
NCCL; on a smaller box (the driver's test box has ONE GPU) the same ranks share cuda:0 and exchange through gloo
(`TimeShard` stages device tensors through the host there), so the sharded orchestration -- boundary messages, halo
rows, warm starts across rank boundaries, the packed statistics all-reduce, the replicated M-step -- runs on the real
kernels either way."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp

from poor_man_gplvm_b200.synthetic import make_dataset

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fit(y, lp0, params, basis_ls, N, K, device, **kw):
    import poor_man_gplvm_b200 as pmg
    m = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=basis_ls, device=device)
    m.params = params.copy()
    res = m.fit_em(y, n_iter=4, log_posterior_init=lp0, m_step_maxiter=25, m_step_tol=-1, **kw)
    dec = m.decode_latent(y, **{k: v for k, v in kw.items() if k == "time_sharded"})
    return res, dec


def _worker(rank, world, port, y, lp0, params, N, K, q, nccl):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank if nccl else 0)
    torch.cuda.set_device(dev)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        T = y.shape[0]
        lo, hi = rank * T // world, (rank + 1) * T // world
        res, dec = _fit(y[lo:hi], lp0[lo:hi], params, 8.0, N, K, dev, time_sharded=True)
        q.put((rank, {"lml": np.array(res["log_marginal_l"]), "tuning": res["tuning"],
                      "post": res["posterior_latent_marg"], "dyn": res["posterior_dynamics_marg"],
                      "dec_lml": dec["log_marginal_final"], "dec_post": dec["posterior_latent_marg"],
                      "pj": np.asarray(dec["p_joint_latent"])}))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_time_sharded_fit_matches_single_gpu(world):
    """world >= 3 exercises interior ranks (two neighbours, both boundary exchanges)."""
    nccl = torch.cuda.device_count() >= world
    import poor_man_gplvm_b200 as pmg
    N, K, T = 30, 96, 6000
    d = make_dataset(T, N, K, seed=21)
    m0 = pmg.PoissonGPLVMJump1D(N, K, tuning_lengthscale=8.0)
    rng = np.random.default_rng(5)
    params = rng.standard_normal((m0.n_basis, N)).astype(np.float32)
    lp0, _ = m0.init_latent_posterior(T, key=3)
    os.environ["PMG_HALO"] = "64"          # several chains per rank even at this small T
    single, sdec = _fit(d["y"], lp0, params, 8.0, N, K, torch.device("cuda", 0))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, d["y"], lp0, params, N, K, q, nccl)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    for r in range(world):
        assert isinstance(out[r], dict), out[r]
    lml1 = np.array(single["log_marginal_l"])
    for r in range(world):
        assert np.max(np.abs(out[r]["lml"] - lml1) / np.abs(lml1)) < 1e-5
        assert np.max(np.abs(out[r]["tuning"] - single["tuning"]) / single["tuning"]) < 1e-3
        assert abs(out[r]["dec_lml"] - sdec["log_marginal_final"]) < 1e-5 * abs(sdec["log_marginal_final"])
        assert np.max(np.abs(out[r]["pj"] - np.asarray(sdec["p_joint_latent"]))) < 1e-5
    for r in range(1, world):
        assert np.array_equal(out[0]["tuning"], out[r]["tuning"])    # replicated M-step: identical on all ranks
    post = np.concatenate([out[r]["post"] for r in range(world)])
    dyn = np.concatenate([out[r]["dyn"] for r in range(world)])
    assert np.max(np.abs(post - single["posterior_latent_marg"])) < 2e-5
    assert np.max(np.abs(dyn - single["posterior_dynamics_marg"])) < 2e-5
    dpost = np.concatenate([out[r]["dec_post"] for r in range(world)])
    assert np.max(np.abs(dpost - sdec["posterior_latent_marg"])) < 2e-5
