"""Forest container and the batched forest ops, served by sm_100a kernels.

Host-side mirror of `src/bark/forest.py` of the reference: same names, argument meaning and array layouts
(`NODE_RECORD_DTYPE` structured arrays in, numpy arrays out), but every op runs on the GPU through the C ABI
in `include/bark_b200.h`.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from enum import Enum

import numpy as np

from . import _lib

# src/bark/forest.py:8-19 -- packed 26-byte records
NODE_RECORD_DTYPE = np.dtype(
    [
        ("is_leaf", np.uint8),
        ("feature_idx", np.uint32),
        ("threshold", np.float32),
        ("left", np.uint32),
        ("right", np.uint32),
        ("parent", np.uint32),
        ("depth", np.uint32),
        ("active", np.uint8),
    ]
)
NODE_LIMIT = 100


class FeatureTypeEnum(Enum):  # src/bark/forest.py:22-25
    Cat = 0
    Int = 1
    Cont = 2


def create_empty_forest(m: int, node_limit: int = NODE_LIMIT) -> np.ndarray:
    """Root-only trees (src/bark/forest.py:114-117).  The reference writes parent = -1 into a uint32
    (0xFFFFFFFF under numpy < 2); written explicitly here so that it also works on numpy >= 2."""
    forest = np.zeros((m, node_limit), dtype=NODE_RECORD_DTYPE)
    forest[:, 0] = (1, 0, 0, 0, 0, 0xFFFFFFFF, 0, 1)
    return forest


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class DeviceForest:
    """Struct-of-arrays copy of `n_nodes` NODE_RECORD_DTYPE records in HBM (lossless, stale fields included)."""

    FIELDS = (("is_leaf", "uint8"), ("active", "uint8"), ("feature", "int32"), ("threshold", "float32"),
              ("left", "int32"), ("right", "int32"), ("parent", "int32"), ("depth", "int32"))

    def __init__(self, shape, device=None):
        torch = _lib.require_cuda()
        self.shape = tuple(int(s) for s in shape)
        self.n_nodes = int(np.prod(self.shape))
        self.device = torch.device(device or "cuda")
        for name, dt in self.FIELDS:
            setattr(self, name, torch.empty(self.n_nodes, dtype=getattr(torch, dt), device=self.device))

    def soa(self) -> _lib.NodesSoA:
        return _lib.NodesSoA(*[getattr(self, n).data_ptr() for n, _ in self.FIELDS])

    @classmethod
    def from_numpy(cls, nodes: np.ndarray, device=None) -> "DeviceForest":
        torch = _lib.require_cuda()
        if nodes.dtype != NODE_RECORD_DTYPE:
            raise TypeError("forest must be a NODE_RECORD_DTYPE structured array")
        self = cls(nodes.shape, device)
        raw = torch.from_numpy(np.ascontiguousarray(nodes).view(np.uint8).reshape(-1)).to(self.device, non_blocking=True)
        _lib.check(_lib.load().bark_nodes_unpack(_ptr(raw), self.n_nodes, self.soa(), _stream()))
        return self

    def pack_into(self, out_bytes, node_offset: int = 0):
        """Pack all records into a device uint8 tensor at record offset `node_offset`."""
        dst = C.c_void_p(out_bytes.data_ptr() + node_offset * NODE_RECORD_DTYPE.itemsize)
        _lib.check(_lib.load().bark_nodes_pack(self.soa(), self.n_nodes, dst, _stream()))

    def to_numpy(self) -> np.ndarray:
        torch = _lib.require_cuda()
        raw = torch.empty(self.n_nodes * NODE_RECORD_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        self.pack_into(raw)
        return raw.cpu().numpy().view(NODE_RECORD_DTYPE).reshape(self.shape)


def _as_device_f64(x, device):
    torch = _lib.require_cuda()
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(device)


def _feat_types_device(feat_types, device):
    torch = _lib.require_cuda()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(feat_types), dtype=np.int32)).to(device)


def traverse_device(dforest: DeviceForest, X_dev, ft_dev):
    """Leaf slot ids (n_forests, N, m) uint32-valued (stored in an int32 tensor) on the device."""
    torch = _lib.require_cuda()
    *lead, m, limit = dforest.shape
    nf = int(np.prod(lead)) if lead else 1
    n, d = X_dev.shape
    out = torch.empty((nf, n, m), dtype=torch.int32, device=X_dev.device)
    _lib.check(_lib.load().bark_traverse(dforest.soa(), nf, m, limit, _ptr(X_dev), n, d, _ptr(ft_dev), _ptr(out),
                                         _stream()))
    return out


def pass_through_forest(nodes: np.ndarray, X: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    """(N, m) uint32 leaf slot of every point in every tree (src/bark/forest.py:58-67).
    Also accepts a batch of forests (..., m, L) -> (..., N, m)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda")
    df = DeviceForest.from_numpy(nodes, dev)
    out = traverse_device(df, _as_device_f64(X, dev), _feat_types_device(feat_types, dev))
    res = out.cpu().numpy().view(np.uint32)
    lead = nodes.shape[:-2]
    return res.reshape(*lead, X.shape[0], nodes.shape[-2])


def pass_through_tree(nodes: np.ndarray, X: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    """(N,) uint32 leaf slots of one tree (src/bark/forest.py:50-55)."""
    return pass_through_forest(nodes.reshape(1, -1), X, feat_types)[:, 0]


def get_leaf_vectors(nodes: np.ndarray, X: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    """(N, B) one-hot indicators of the sorted non-empty leaves of ONE tree (src/bark/forest.py:70-75).
    The traversal runs on the GPU; the tiny (B <= ~5) unique/one-hot step is done on the returned ids.
    (The sampler itself never materialises these: it keeps leaf bitsets on the device.)"""
    torch = _lib.require_cuda()
    dev = torch.device("cuda")
    df = DeviceForest.from_numpy(nodes.reshape(1, -1), dev)
    ids = traverse_device(df, _as_device_f64(X, dev), _feat_types_device(feat_types, dev))[0, :, 0]
    present = torch.unique(ids)
    return (ids[:, None] == present[None, :]).to(torch.float64).cpu().numpy()


def gram_counts_device(leaves_a, leaves_b, slots=None):
    """Exact int32 co-occurrence counts (batch, Na, Nb) from leaf ids (batch, Na, m), (batch, Nb, m)."""
    return gram_umma_device(leaves_a, leaves_b, slots=slots)[0]


def forest_slots(nodes: np.ndarray) -> int:
    """Upper bound on (leaf slot id + 1) over a batch of forests: highest active slot + 1."""
    act = np.flatnonzero(nodes["active"].reshape(-1, nodes.shape[-1]).any(axis=0))
    return int(act.max()) + 1 if act.size else 1


def gram_umma_device(leaves_a, leaves_b, slots=None, want_counts=True, scale=None, noise=None, jitter=1e-6):
    """Tensor-core Gram (int8 one-hot tcgen05 GEMM): returns (counts int32 or None, K f64 or None).
    K is produced when `scale` is given; the diagonal (jitter + noise) is added when `noise` is given."""
    torch = _lib.require_cuda()
    lib = _lib.load()
    b, na, m = leaves_a.shape
    nb = leaves_b.shape[1]
    if slots is None:  # ids must be < slots: one small device reduction + host read
        slots = int(max(int(leaves_a.max().item()), int(leaves_b.max().item())) + 1) if leaves_a.numel() else 1
    dev = leaves_a.device
    counts = torch.empty((b, na, nb), dtype=torch.int32, device=dev) if want_counts else None
    K = torch.empty((b, na, nb), dtype=torch.float64, device=dev) if scale is not None else None
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = int(lib.bark_gram_workspace_bytes(b, na, nb, m, slots))
    ws = torch.empty(max(nbytes, 8), dtype=torch.uint8, device=dev)
    _lib.check(lib.bark_gram_umma(_ptr(leaves_a), _ptr(leaves_b), b, na, nb, m, slots, _ptr(counts), _ptr(K), _ptr(scale),
                                  _ptr(noise), float(jitter), int(noise is not None), _ptr(status), _ptr(ws), _stream()))
    st = int(status.item())
    if st & 1:
        raise _lib.BarkError("leaf id >= slots in bark_gram_umma")
    if st & 2:
        raise _lib.BarkError("device pipeline wait timed out in bark_gram_umma (internal error)")
    return counts, K


def gram_to_kernel_device(counts, m, scale, noise=None, jitter=1e-6):
    """K = scale*((1/m)*counts) [+ (jitter+noise) I], FP64, same rounding as the reference."""
    torch = _lib.require_cuda()
    b, na, nb = counts.shape
    K = torch.empty((b, na, nb), dtype=torch.float64, device=counts.device)
    add = noise is not None
    _lib.check(_lib.load().bark_gram_to_kernel(_ptr(counts), b, na, nb, m, _ptr(scale), _ptr(noise) if add else C.c_void_p(0),
                                               float(jitter), int(add), _ptr(K), _stream()))
    return K


def forest_gram_counts(nodes: np.ndarray, x1: np.ndarray, x2: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    """Integer leaf co-occurrence counts (..., N, M) int32 -- the bit-exact contract of the Gram."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda")
    df = DeviceForest.from_numpy(nodes, dev)
    ft = _feat_types_device(feat_types, dev)
    la = traverse_device(df, _as_device_f64(x1, dev), ft)
    lb = la if x2 is x1 else traverse_device(df, _as_device_f64(x2, dev), ft)
    cnt = gram_counts_device(la, lb, slots=forest_slots(nodes)).cpu().numpy()
    return cnt.reshape(*nodes.shape[:-2], x1.shape[0], x2.shape[0])


def forest_gram_matrix(nodes: np.ndarray, x1: np.ndarray, x2: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    """K0 = (1/m) * count, (N, M) f64 (src/bark/forest.py:78-89)."""
    return batched_forest_gram_matrix(nodes.reshape(1, *nodes.shape), x1, x2, feat_types)[0]


def batched_forest_gram_matrix(nodes: np.ndarray, x1: np.ndarray, x2: np.ndarray, feat_types: np.ndarray) -> np.ndarray:
    """(S, N, M) f64 (src/bark/forest.py:92-98)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda")
    m = nodes.shape[-2]
    df = DeviceForest.from_numpy(nodes, dev)
    ft = _feat_types_device(feat_types, dev)
    la = traverse_device(df, _as_device_f64(x1, dev), ft)
    lb = la if x2 is x1 else traverse_device(df, _as_device_f64(x2, dev), ft)
    ones = torch.ones(la.shape[0], dtype=torch.float64, device=dev)
    _, K0 = gram_umma_device(la, lb, slots=forest_slots(nodes), want_counts=False, scale=ones)  # fused FP64 epilogue
    return K0.cpu().numpy()


def batched_forest_gram_matrix_no_null(nodes: np.ndarray, x1, x2, feat_types) -> np.ndarray:
    """Gram after removing root-only trees (src/bark/forest.py:101-111); consumed by the (host-side,
    out-of-scope) acquisition MIP."""
    sim = batched_forest_gram_matrix(nodes, x1, x2, feat_types)
    m = nodes.shape[-2]
    n_null = np.sum(nodes[:, :, 0]["is_leaf"], axis=-1)[:, None, None]
    return (sim - n_null / m) * (m / np.maximum(m - n_null, 1))
