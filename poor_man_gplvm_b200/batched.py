"""Batched callers of the hot path.

The reference calls ``decode_latent`` / ``decode_latent_naive_bayes`` hundreds of times on the same spike matrix:
``model_selection_helper.get_downsampled_lml`` (:243-260, n_repeat latent masks), ``get_lml_test_history``
(:424-445, one call per saved tuning) and ``test.shuffle_and_decode`` / ``test_one_model`` (test.py:27-63, n_shuffle
circular shuffles).  Every such call re-uploads the spikes, re-derives the fp16 counts and the lgamma row term, and --
for the ones that only read a log marginal -- runs a backward pass nobody looks at.

Here the recording is prepared once (:class:`DecodeSession`: spikes on the device, emission operands, E-step
buffers) and each variant costs one emission GEMM + one forward scan (``forward_only``) or one full decode.  Same
function names, arguments and result layout as the reference; masks / shuffles are drawn with NumPy generators
(``jax.random.choice(replace=False)`` and the unseeded ``np.random`` of the reference are not reproducible streams),
or passed in explicitly (``latent_masks=``, ``shifts=``) for run-for-run comparisons.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .core import _seed_from_key, _unwrap_tsd
from .estep import EStep


class DecodeSession:
    """One recording, many decodes: the spikes live on the device, the count-derived emission operands and the scan
    buffers are built once per neuron mask."""

    def __init__(self, model, y, ma_neuron=None, likelihood_scale=1.0, hyperparam={}):
        """ma_neuron: None, [N] or [T,N] (fixed for the session: the count-derived operands depend on it)."""
        self.model = model
        y, self.t_l = _unwrap_tsd(y)
        with _guard(model):
            self.y = model._dev(y)
            self.T = self.y.shape[0]
            _, _, _, _, self.op = model._transition_pack(dict(hyperparam))
            ma_n, _ = model._masks(ma_neuron, None, self.T)
            self.es = EStep(self.y, self.op, ma_n, None, float(likelihood_scale))

    def log_marginal(self, tuning=None, ma_latent=None):
        """log_marginal_final of decode_latent (reference core.py:485) from the forward filter alone."""
        m = self.model
        with _guard(m):
            es = self.es
            es.ma_latent = m._dev(m.ma_latent_default if ma_latent is None else ma_latent)
            res = es.run(m._dev(m.tuning if tuning is None else tuning), forward_only=True)
            return float(res.log_marginal)

    def naive_bayes_total(self, tuning=None, ma_latent=None):
        """log_marginal_total of decode_latent_naive_bayes (reference core.py:520)."""
        m = self.model
        with _guard(m):
            es = self.es
            ma_l = m._dev(m.ma_latent_default if ma_latent is None else ma_latent)
            ll = es.em.loglik(m._dev(m.tuning if tuning is None else tuning), ma_l, 1.0, out=es.ll)
            _, lml = ops.naive_bayes_normalize(ll, inplace=True)
            return float(lml.sum(dtype=torch.float64).item())


def _guard(model):
    import contextlib
    dev = model.device
    return torch.cuda.device(dev) if dev.type == "cuda" else contextlib.nullcontext()


# ------------------------------------------------------------------------------------------------------------
# model_selection_helper.py
# ------------------------------------------------------------------------------------------------------------
def draw_latent_masks(n_latent_bin, downsample_frac, n_repeat, key=4):
    """n_repeat 0/1 masks with int(K * frac) ones each (reference model_selection_helper.py:249-254; NumPy stream)."""
    rng = np.random.default_rng(_seed_from_key(key))
    n_sel = int(n_latent_bin * downsample_frac)
    masks = np.zeros((n_repeat, n_latent_bin), dtype=np.float32)
    for i in range(n_repeat):
        masks[i, rng.choice(n_latent_bin, size=n_sel, replace=False)] = 1.0
    return masks


def get_downsampled_lml(model_fit, y_test, downsample_frac=0.2, n_repeat=10, key=4, latent_masks=None, **kwargs):
    """reference model_selection_helper.py:243-260: mean / std over n_repeat random latent masks of the log marginal
    of ``decode_latent(y_test, ma_latent=mask)``.  kwargs: tuning, hyperparam, ma_neuron, likelihood_scale (as
    ``decode_latent``).  One emission GEMM + one forward scan per mask on a shared :class:`DecodeSession`."""
    if latent_masks is None:
        latent_masks = draw_latent_masks(model_fit.n_latent_bin, downsample_frac, n_repeat, key)
    ses = DecodeSession(model_fit, y_test, ma_neuron=kwargs.get("ma_neuron"),
                        likelihood_scale=kwargs.get("likelihood_scale", 1.0), hyperparam=kwargs.get("hyperparam", {}))
    lml_l = [ses.log_marginal(tuning=kwargs.get("tuning"), ma_latent=mask) for mask in np.asarray(latent_masks)]
    return {'value': np.mean(lml_l), 'std': np.std(lml_l), 'lml_l': lml_l}


def get_lml_test_history(y_test, model, tuning_saved, do_nb=True, ma_temporal=None):
    """reference model_selection_helper.py:424-445: test log marginal under every saved tuning (naive Bayes total or
    the smoother's log marginal); ``ma_temporal`` [T] expands to a [T,N] neuron mask."""
    ma_neuron = None
    if ma_temporal is not None:
        y_arr, _ = _unwrap_tsd(y_test)
        ma_neuron = np.ones((1, np.shape(y_arr)[1]), np.float32) * np.asarray(ma_temporal, np.float32)[:, None]
    ses = DecodeSession(model, y_test, ma_neuron=ma_neuron)
    fn = ses.naive_bayes_total if do_nb else ses.log_marginal
    return np.array([fn(tuning=tun) for tun in tuning_saved])


# ------------------------------------------------------------------------------------------------------------
# test.py (shuffle tests; not pytest tests)
# ------------------------------------------------------------------------------------------------------------
def draw_shifts(n_time, n_neuron, n_shuffle, seed=None):
    """[n_shuffle, n_neuron] circular shifts in [0, n_time) (the reference draws them with the global np.random)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_time, size=(n_shuffle, n_neuron))


def _roll_columns(y_dev, shift_dev):
    """y_shuffled[t, n] = y[(t - shift[n]) mod T, n]  (np.roll per neuron, reference test.py:22-23)."""
    T = y_dev.shape[0]
    idx = (torch.arange(T, device=y_dev.device).unsqueeze(1) - shift_dev.unsqueeze(0)) % T
    return torch.gather(y_dev, 0, idx)


def circular_shuffle_data(spk_tsdf, n_shuffle=100, ep=None, shifts=None, seed=None, device=None):
    """reference test.py:10-24: every neuron circularly shifted by its own random offset; yields n_shuffle arrays
    (device tensors when the input is one or ``device`` is given, else NumPy)."""
    if ep is not None:
        spk_tsdf = spk_tsdf.restrict(ep)                 # pynapple TsdFrame
    y, _ = _unwrap_tsd(spk_tsdf)
    on_dev = isinstance(y, torch.Tensor) or device is not None
    if on_dev and not isinstance(y, torch.Tensor):
        y = torch.as_tensor(np.asarray(y, dtype=np.float32)).to(device)
    n_time, n_neuron = y.shape
    if shifts is None:
        shifts = draw_shifts(n_time, n_neuron, n_shuffle, seed)
    for i in range(n_shuffle):
        if on_dev:
            yield _roll_columns(y, torch.as_tensor(np.asarray(shifts[i]), device=y.device))
        else:
            out = np.empty_like(np.asarray(y))
            for j in range(n_neuron):
                out[:, j] = np.roll(np.asarray(y)[:, j], int(shifts[i][j]))
            yield out


def shuffle_and_decode(model, spk_tsdf, n_time_per_chunk=10000, dt_l=1, n_shuffle=100, ep=None,
                       decoder_type='naive_bayes', shifts=None, seed=None, keys=None):
    """reference test.py:27-45: decode n_shuffle circular shuffles; every result key stacked over the shuffles.
    The spikes are uploaded once and shuffled on the device.  keys: optional subset of result keys to keep (the
    reference stacks all of them: n_shuffle x T x K floats for the posteriors)."""
    if decoder_type not in ('naive_bayes', 'dynamics'):
        raise ValueError(f"decoder_type {decoder_type} not supported")
    if ep is not None:
        spk_tsdf = spk_tsdf.restrict(ep)
    y, _ = _unwrap_tsd(spk_tsdf)
    with _guard(model):
        y_dev = model._dev(y)
        out = None
        for y_sh in circular_shuffle_data(y_dev, n_shuffle=n_shuffle, shifts=shifts, seed=seed):
            if decoder_type == 'naive_bayes':
                res = model.decode_latent_naive_bayes(y_sh, n_time_per_chunk=n_time_per_chunk, dt_l=dt_l)
            else:
                res = model.decode_latent(y_sh, n_time_per_chunk=n_time_per_chunk)
            if out is None:
                out = {k: [] for k in res if keys is None or k in keys}
            for k in out:
                out[k].append(np.asarray(res[k]))
    return {k: np.array(v) for k, v in out.items()}


def test_one_model(y_true, model_fit, n_shuffle=100, decoder_type='naive_bayes', sig_key=None, seed=None):
    """reference test.py:48-66: per-bin significance of the decode against the 97.5 % quantile over shuffles."""
    y_val, y_t = _unwrap_tsd(y_true)
    if sig_key is None:
        sig_key = 'log_marginal_l' if decoder_type == 'naive_bayes' else 'log_one_step_predictive_marginals_all'
    if decoder_type == 'naive_bayes':
        res_true = model_fit.decode_latent_naive_bayes(y_val)
    elif decoder_type == 'dynamics':
        res_true = model_fit.decode_latent(y_val)
    else:
        raise ValueError(f"decoder_type {decoder_type} not supported")
    res_shuffle = shuffle_and_decode(model_fit, y_val, n_time_per_chunk=10000, dt_l=1, n_shuffle=n_shuffle, ep=None,
                                     decoder_type=decoder_type, seed=seed)
    log_marg_thresh = np.quantile(res_shuffle[sig_key], 0.975, axis=0)
    is_sig = np.asarray(res_true[sig_key]) > log_marg_thresh
    is_sig_tsd = is_sig
    if y_t is not None:
        try:
            import pynapple as nap
            is_sig_tsd = nap.Tsd(d=is_sig, t=y_t)
        except Exception:
            pass
    return {'decode_res_true': res_true, 'decode_res_shuffle': res_shuffle, 'log_marg_thresh': log_marg_thresh,
            'is_sig_tsd': is_sig_tsd}


test_one_model.__test__ = False          # a shuffle test of a model, not a pytest test


def compute_entropy(logp_l, axis=(-1, -2)):
    """reference test.py:68-79: -sum p log p over `axis`."""
    logp_l = np.asarray(logp_l)
    with np.errstate(invalid="ignore"):
        term = np.where(np.isneginf(logp_l), 0.0, np.exp(logp_l) * logp_l)      # 0 log 0 = 0 (our logs may be -inf)
    return -np.sum(term, axis=axis)
