"""On-disk format for posterior samples.  The reference has none -- `_dumps` / `loads` of
`src/bofire_mixed/surrogates/bark.py:96-100` are `pass` -- so this is the natural one: a single `.npz` holding
the sampler's return value byte for byte (AoS `NODE_RECORD_DTYPE` records as raw bytes, so the packed 26-byte
layout survives), the hyper-parameter samples, the training data and the sampler parameters.  A fit can be
resumed from it with `warmup_steps=0`, exactly like the reference's warm start across `tell()` calls
(`surrogates/bark.py:131-141`)."""
from __future__ import annotations

import json

import numpy as np

from .forest import NODE_RECORD_DTYPE
from .sampler import BARKTrainParams

FORMAT_VERSION = 1
_PARAM_FIELDS = ("warmup_steps", "num_samples", "steps_per_sample", "num_chains", "alpha", "beta", "verbose",
                 "use_softplus_transform", "sample_scale", "gamma_prior_shape", "gamma_prior_rate")


def save_samples(path, model, data=None, params: BARKTrainParams | None = None, extra: dict | None = None) -> None:
    """model = (node_samples (..., m, L) NODE_RECORD_DTYPE, noise (...), scale (...)); data = (X, y) or None."""
    forest, noise, scale = model
    forest = np.ascontiguousarray(forest)
    if forest.dtype != NODE_RECORD_DTYPE:
        raise TypeError("forest samples must have NODE_RECORD_DTYPE")
    meta = {"format_version": FORMAT_VERSION, "forest_shape": list(forest.shape), "extra": extra or {}}
    if params is not None:
        meta["params"] = {k: getattr(params, k) for k in _PARAM_FIELDS}
        meta["params"]["proposal_weights"] = [float(w) for w in params.proposal_weights]
    arrays = {"forest_bytes": forest.view(np.uint8).reshape(-1), "noise": np.asarray(noise, dtype=np.float64),
              "scale": np.asarray(scale, dtype=np.float64), "meta": np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)}
    if data is not None:
        arrays["train_x"] = np.asarray(data[0], dtype=np.float64)
        arrays["train_y"] = np.asarray(data[1], dtype=np.float64)
    with open(path, "wb") as fh:
        np.savez_compressed(fh, **arrays)


def load_samples(path):
    """-> dict(model=(forest, noise, scale), data=(X, y) | None, params=BARKTrainParams | None, extra=dict)."""
    with np.load(path) as z:
        meta = json.loads(bytes(z["meta"]).decode())
        if meta.get("format_version") != FORMAT_VERSION:
            raise ValueError(f"unsupported checkpoint version {meta.get('format_version')}")
        forest = np.frombuffer(z["forest_bytes"].tobytes(), dtype=NODE_RECORD_DTYPE).reshape(meta["forest_shape"]).copy()
        model = (forest, z["noise"].copy(), z["scale"].copy())
        data = (z["train_x"].copy(), z["train_y"].copy()) if "train_x" in z.files else None
    params = None
    if "params" in meta:
        kw = dict(meta["params"])
        kw["proposal_weights"] = np.asarray(kw["proposal_weights"], dtype=np.float64)
        params = BARKTrainParams(**kw)
    return {"model": model, "data": data, "params": params, "extra": meta.get("extra", {})}
