"""The pieces of `oracle/jaxshim` that RESTATE third-party code (optax.adam, jax.scipy.special, jax.nn) are
builder-written, so a shared misreading of optax / JAX would pass the golden-fixture tests unnoticed (VERDICT round 1,
"pin caveat").  These tests hold them to INDEPENDENT implementations of the same published definitions that ship in
this image: torch.optim.Adam (Kingma & Ba with bias correction; identical to optax.adam with eps_root = 0),
scipy.special (logsumexp, xlogy, gammaln) and torch.nn.functional (softplus, sigmoid, softmax).  CPU only.

The shim fixes its float width at import (JAXSHIM_X64, like JAX's x64 switch) and sets torch's default dtype, so each
check runs in its own interpreter."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "oracle", "jaxshim")

_ADAM = r'''
import json, sys
import numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[2])
import jax, optax
import jax.numpy as jnp
assert str(getattr(jax, "__version__", "")).endswith("shim")
dtype = np.float64 if sys.argv[3] == "1" else np.float32
rng = np.random.default_rng(0)
p0 = rng.standard_normal((7, 5)).astype(dtype)
grads = [rng.standard_normal((7, 5)).astype(dtype) * (1.0 + 0.1 * i) for i in range(60)]
lr = 0.01
tp = torch.nn.Parameter(torch.from_numpy(p0.copy()))
opt = torch.optim.Adam([tp], lr=lr, betas=(0.9, 0.999), eps=1e-8)
for g in grads:
    tp.grad = torch.from_numpy(g.copy())
    opt.step()
want = tp.detach().numpy().astype(np.float64)
tx = optax.adam(lr)
params = jnp.asarray(p0.copy())
state = tx.init(params)
for g in grads:
    upd, state = tx.update(jnp.asarray(g.copy()), state, params)
    params = optax.apply_updates(params, upd)
got_shim = np.asarray(params)
f = dtype
p, mu, nu, count = p0.copy(), np.zeros_like(p0), np.zeros_like(p0), 0
for g in grads:                                  # the recursion of oracle.ref_numpy.adam_run's loop body
    mu = f(0.9) * mu + f(1 - 0.9) * g
    nu = f(0.999) * nu + f(1 - 0.999) * (g * g)
    count += 1
    p = p + f(-lr) * ((mu / f(1 - f(0.9) ** count)) / (np.sqrt(nu / f(1 - f(0.999) ** count)) + f(1e-8)))
from oracle import ref_numpy as ref
import inspect
src = inspect.getsource(ref.adam_run)
print(json.dumps({"shim_dtype": str(got_shim.dtype), "shim": float(np.max(np.abs(got_shim - want))),
                  "oracle": float(np.max(np.abs(p - want))), "count": int(np.asarray(state[0].count)),
                  "oracle_has_recursion": ("mu_hat / (np.sqrt(nu_hat) + f(eps))" in src)}))
'''

_SPECIAL = r'''
import json, sys
import numpy as np, torch, scipy.special as sp
sys.path.insert(0, sys.argv[1])
import jax, jax.numpy as jnp, jax.scipy.special as jss, jax.nn as jnn
rng = np.random.default_rng(1)
out = {}
a = rng.standard_normal((6, 9)) * 30
a[2, 3] = -np.inf
out["logsumexp"] = max(float(np.max(np.abs(np.asarray(jss.logsumexp(jnp.asarray(a), axis=ax)) - sp.logsumexp(a, axis=ax))
                                    / np.abs(sp.logsumexp(a, axis=ax)))) for ax in (None, 0, 1))
x = np.array([0.0, 0.0, 2.5, 1e-30, 7.0]); y = np.array([0.0, 3.0, 1e-20, 5.0, 0.0])
with np.errstate(divide="ignore"):
    out["xlogy_equal"] = bool(np.array_equal(np.asarray(jss.xlogy(jnp.asarray(x), jnp.asarray(y))), sp.xlogy(x, y)))
z = np.array([1.0, 2.0, 3.5, 17.0, 1e-3, 250.0])
out["gammaln"] = float(np.max(np.abs(np.asarray(jss.gammaln(jnp.asarray(z))) - sp.gammaln(z)) / np.abs(sp.gammaln(z) + 1e-300)))
t = rng.standard_normal(50) * 20
tt = torch.from_numpy(t)
out["softplus"] = float(np.max(np.abs(np.asarray(jnn.softplus(jnp.asarray(t))) - np.logaddexp(t, 0.0)) / np.logaddexp(t, 0.0)))
out["softplus_torch"] = float(np.max(np.abs(np.asarray(jnn.softplus(jnp.asarray(t)))
                                            - torch.nn.functional.softplus(tt, threshold=1e9).numpy()) / np.logaddexp(t, 0.0)))
out["sigmoid"] = float(np.max(np.abs(np.asarray(jnn.sigmoid(jnp.asarray(t))) - torch.sigmoid(tt).numpy())))
m = rng.standard_normal((4, 7))
out["softmax"] = float(np.max(np.abs(np.asarray(jnn.softmax(jnp.asarray(m), axis=1)) - torch.softmax(torch.from_numpy(m), dim=1).numpy())))
print(json.dumps(out))
'''


def _run(code, x64):
    env = dict(os.environ, JAXSHIM_X64="1" if x64 else "0")
    r = subprocess.run([sys.executable, "-c", code, SHIM, ROOT, "1" if x64 else "0"], capture_output=True, text=True,
                       env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("x64", [True, False])
def test_optax_adam_shim_and_oracle_adam_follow_torch_adam(x64):
    """60 Adam steps on a fixed gradient sequence: the shim's optax.adam, the update recursion of
    oracle.ref_numpy.adam_run (fit_tuning_helper.py:124-196 restated) and torch.optim.Adam give the same iterates."""
    res = _run(_ADAM, x64)
    tol = 1e-12 if x64 else 3e-6
    assert res["shim_dtype"] == ("float64" if x64 else "float32")
    assert res["count"] == 60
    assert res["shim"] < tol and res["oracle"] < tol, res
    assert res["oracle_has_recursion"]


def test_special_functions_of_the_shim_match_scipy_and_torch():
    res = _run(_SPECIAL, True)
    assert res["xlogy_equal"]
    for k in ("logsumexp", "gammaln", "softplus", "softplus_torch", "sigmoid", "softmax"):
        assert res[k] < 1e-12, (k, res)
