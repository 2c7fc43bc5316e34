"""bark_b200 -- B200-native (sm_100a) implementation of the BARK tree-kernel hot path.

Drop-in for the reference's `run_bark_sampler` / `forest_predict` / forest ops; every op runs in hand-written
CUDA kernels behind the C ABI of `include/bark_b200.h`.  There is no CPU fallback."""
from ._build import build  # noqa: F401
from ._lib import BarkError, load  # noqa: F401
from .forest import (NODE_RECORD_DTYPE, FeatureTypeEnum, batched_forest_gram_matrix,  # noqa: F401
                     batched_forest_gram_matrix_no_null, create_empty_forest, forest_gram_counts, forest_gram_matrix,
                     get_leaf_vectors, pass_through_forest, pass_through_tree)
from .mll import forest_mll  # noqa: F401
from .predict import BARKModel, PosteriorState, forest_predict, mixture_of_gaussians_as_normal  # noqa: F401
from .sampler import BARKTrainParams, BARKTrainParamsNumba, ChainState, run_bark_sampler  # noqa: F401
from .acquisition import gp_sample_inverses  # noqa: F401
from .checkpoint import load_samples, save_samples  # noqa: F401
from .prior import sample_forest_prior, sample_forest_prior_device, sample_noise_prior  # noqa: F401
from .tree_kernel import TreeAgreementKernel  # noqa: F401
from .surrogate import BARKPriorSurrogate, BARKSurrogate, Standardize  # noqa: F401
