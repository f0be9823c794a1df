from . import special  # noqa: F401
from . import stats  # noqa: F401
