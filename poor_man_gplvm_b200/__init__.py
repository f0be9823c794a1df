"""poor_man_gplvm_b200 — B200-native EM hot path of PoissonGPLVMJump1D.

Drop-in surface (same names as the reference package ``poor_man_gplvm``):
``PoissonGPLVMJump1D`` with ``fit_em`` / ``decode_latent`` /
``decode_latent_naive_bayes`` / ``tuning`` and the reference's result keys.
"""
from .core import PoissonGPLVMJump1D, compute_transition_posterior_prob  # noqa: F401
from .gp_kernel import create_transition_prob_1d, generate_basis  # noqa: F401
from .families import GaussianGPLVM1D, GaussianGPLVMJump1D, PoissonGPLVM1D  # noqa: F401
from . import batched, model_selection_helper  # noqa: F401

__all__ = ["PoissonGPLVMJump1D", "GaussianGPLVMJump1D", "PoissonGPLVM1D", "GaussianGPLVM1D",
           "compute_transition_posterior_prob", "create_transition_prob_1d", "generate_basis"]
__version__ = "0.1.0"
