"""Pre-computed-top-k SpGEMM operator, generation 2 surface (reference file `spgemmfunction_v2`, :18-188).

The reference's v2 is its "optimized" operator (spgemmfunction.py) under the plain names
MaxKSpGEMMFunction / maxk_spgemm / MaxKSpmmWrapper: top-k values + indices computed by the caller,
forward = SpGEMM then / in_degrees (:62-76), backward = grad / out_degrees then SSpMM over the CSC arrays
(:93-104), 11 inputs -> 11 gradients with only topk_values receiving one (:106).  Both divisions are fused into
our kernels; every `assert ... REQUIRED` of the reference (:47-50, :153, :181) raises RuntimeError here.
"""
from spgemmfunction import (OptimizedMaxKSpGEMMFunction as MaxKSpGEMMFunction,     # noqa: F401
                            OptimizedMaxKSpmmWrapper as MaxKSpmmWrapper,             # noqa: F401
                            optimized_maxk_spgemm as maxk_spgemm)                    # noqa: F401

MAXK_KERNELS_AVAILABLE = True
