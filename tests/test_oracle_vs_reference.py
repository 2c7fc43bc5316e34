"""Oracle restatement vs the UNMODIFIED reference imported from /root/reference (build container only;
skipped on the GPU box, where the committed golden fixtures take over)."""
import warnings

import numpy as np
import pytest

from oracle import bark_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not mounted")
warnings.filterwarnings("ignore")


@pytest.fixture(scope="module")
def ref():
    ref_shim.install()
    import bark.forest as RF
    from bark.fitting import bark_sampler as RS
    from bark.fitting import quick_inverse as RQ
    from bark.fitting import tree_proposals as RT
    return dict(RF=RF, RS=RS, RQ=RQ, RT=RT)


def test_dtype_identical(ref):
    assert O.NODE_RECORD_DTYPE == ref["RF"].NODE_RECORD_DTYPE
    assert O.create_empty_forest(3).tobytes() == ref_shim.empty_forest(3).tobytes()


def test_seeded_sampler_identical(ref):
    X, y, bounds, ft, _ = O.synthetic_problem(45, dim=2, cat_dim=2, num_cat=3, m_true=6, seed=11)
    C, m = 2, 7
    p = O.BARKTrainParams(warmup_steps=6, num_samples=2, steps_per_sample=2, num_chains=C)
    pr = ref["RS"].BARKTrainParamsNumba(6, 2, 2, C, 0.95, 2.0, np.array([.25, .25, .5]), False, True, False, 1.5, 5.0)
    f0 = np.tile(O.create_empty_forest(m), (C, 1, 1))
    O.seed_numba(5)
    a = O.run_bark_sampler((f0.copy(), np.full(C, 0.1), np.full(C, 1.0)), (X, y), bounds, ft, p)
    O.seed_numba(5)
    b = ref["RS"]._run_bark_sampler_multichain(f0.copy(), np.full(C, 0.1), np.full(C, 1.0), X, y, bounds, ft, pr)
    assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    forest = a[0][0, -1]
    assert np.array_equal(O.pass_through_forest(forest, X, ft), ref["RF"].pass_through_forest(forest, X, ft))
    assert np.array_equal(O.forest_gram_matrix(forest, X, X, ft), ref["RF"].forest_gram_matrix(forest, X, X, ft))
