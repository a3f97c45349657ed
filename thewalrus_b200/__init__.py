"""walrus_b200 — B200-native (sm_100a CUDA) exponential-sum hot path of The Walrus.

Drop-in for the reference's matrix functions on that path:
``hafnian``, ``loop_hafnian``, ``hafnian_repeated``, ``perm``, ``tor``, ``ltor``, ``loop_hafnian_batch`` and the
batched GBS-probability front end.  See DESIGN.md for scope and INTEGRATION.md for the binding.
"""
from ._hafnian import (  # noqa: F401
    _haf,
    find_kept_edges,
    hafnian,
    hafnian_batch,
    hafnian_repeated,
    input_validation,
    loop_hafnian,
    matched_reps,
    reduction,
)
from .loop_hafnian_batch import loop_hafnian_batch  # noqa: F401
from .loop_hafnian_batch_gamma import loop_hafnian_batch_gamma  # noqa: F401
from ._permanent import (  # noqa: F401
    brs,
    fock_prob,
    fock_threshold_prob,
    perm,
    perm_bbfg,
    perm_ryser,
    permanent_repeated,
    ubrs,
)
from ._torontonian import (  # noqa: F401
    ltor,
    numba_ltor,
    numba_tor,
    numba_vac_prob,
    rec_ltorontonian,
    rec_torontonian,
    threshold_detection_prob,
    tor,
    tor_input_checks,
)
from ._montrealer import lmtl, mtl  # noqa: F401
from . import quantum  # noqa: F401  (first: quantum re-exports moments at its end, moments imports quantum)
from . import moments, samples  # noqa: F401
from .quantum import (  # noqa: F401
    density_matrix,
    density_matrix_element,
    probabilities,
    probabilities_batch,
    pure_state_amplitude,
    state_vector,
)

__version__ = "0.1.0"


def version():
    """Version string (thewalrus/__init__.py:166-175)."""
    return __version__
