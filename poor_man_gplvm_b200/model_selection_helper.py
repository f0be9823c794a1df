"""Same module name as the reference's ``poor_man_gplvm/model_selection_helper.py`` for the two metrics that sit on the
hot path (:243-260 ``get_downsampled_lml``, :424-445 ``get_lml_test_history``); implementation in ``batched.py``.
The grid-search drivers of that file are plain Python over ``fit_em`` and work unchanged with this model class."""
from .batched import draw_latent_masks, get_downsampled_lml, get_lml_test_history  # noqa: F401
