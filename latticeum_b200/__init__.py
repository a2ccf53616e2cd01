"""latticeum_b200 -- B200-native Ajtai commitment engine for Latticeum's LatticeFold prover.

Only the one hot path: iCRT -> balanced gadget decomposition -> CRT -> kappa x n Ajtai mat-vec in CRT form over
the Goldilocks ring Z_q[X]/(X^24 - X^12 + 1).  The product is the C-ABI shared library
(include/lattice_ajtai.h, built by latticeum_b200.build); this package is its host-side mirror of the
reference's API.  There is no CPU fallback.
"""
from .scheme import (  # noqa: F401
    AjtaiCommitmentScheme,
    Commitment,
    CommitPipeline,
    CommitmentError,
    DecompositionParams,
    DigitOverflow,
    EngineError,
    FoldStep,
    GoldiLocksDP,
    KAPPA,
    LFDecompositionProver,
    LFFoldingProver,
    N,
    W_SIZE,
    Witness,
    WrongAjtaiMatrixDimensions,
    WrongCommitmentLength,
    WrongWitnessLength,
    from_mont,
    gadget_decompose,
    gadget_recompose,
    digits_to_fq,
    get_fhat,
    get_fhat_from_digits,
    ntt_from_scalar,
    ntt_negacyclic,
    pinned_empty,
    to_mont,
)

__all__ = [
    "AjtaiCommitmentScheme", "Commitment", "ntt_negacyclic", "CommitPipeline", "pinned_empty", "CommitmentError", "DecompositionParams", "DigitOverflow", "EngineError",
    "GoldiLocksDP", "KAPPA", "LFDecompositionProver", "LFFoldingProver", "gadget_decompose", "gadget_recompose", "N", "W_SIZE", "Witness", "WrongAjtaiMatrixDimensions",
    "WrongCommitmentLength", "WrongWitnessLength", "from_mont", "get_fhat", "get_fhat_from_digits", "digits_to_fq", "FoldStep", "ntt_from_scalar", "to_mont",
]
