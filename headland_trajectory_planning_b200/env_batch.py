"""Pack (environment, vehicle, guide) triples into ``HlEnvHost`` records and upload
them to HBM (``hl_env_upload``).  One ``EnvBatch`` holds the static geometry of many
independent scenarios; kernels index it by ``env_id``."""
import ctypes as C

import numpy as np

from . import _lib
from .geometry_host import CAPSULE_VERTS


def _c64(a, shape=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


class EnvRecord:
    """Host arrays of one environment (kept alive while the ctypes struct is used)."""

    def __init__(self, obstacle_quads, field_poly, body_ext, aux_exts=None, seg_xy=None,
                 seg_polys=None, seg_len=None, crit_xy=None, guide=None, default_search_length=1.5):
        self.obs = _c64(obstacle_quads, (-1, 4, 2))
        self.field = _c64(field_poly, (-1, 2))
        self.body_ext = _c64(body_ext, (4,))
        self.aux = _c64(aux_exts if aux_exts is not None else np.zeros((0, 4)), (-1, 4))
        self.seg_xy = _c64(seg_xy if seg_xy is not None else np.zeros((0, 2, 2)), (-1, 2, 2))
        self.seg_polys = _c64(seg_polys if seg_polys is not None else np.zeros((0, CAPSULE_VERTS, 2)),
                              (-1, CAPSULE_VERTS, 2))
        self.seg_len = _c64(seg_len if seg_len is not None else np.zeros((0,)), (-1,))
        self.crit = _c64(crit_xy if crit_xy is not None else np.zeros((0, 2)), (-1, 2))
        self.guide = _c64(guide if guide is not None else np.zeros((0, 4)), (-1, 4))
        self.default_search_length = float(default_search_length)
        assert len(self.seg_polys) == len(self.seg_xy) == len(self.seg_len)

    def as_struct(self):
        h = _lib.HlEnvHost()
        h.n_obs = len(self.obs); h.obs_xy = _lib.dptr(self.obs)
        h.n_field = len(self.field); h.field_xy = _lib.dptr(self.field)
        h.n_seg = len(self.seg_xy); h.seg_xy = _lib.dptr(self.seg_xy)
        h.seg_poly = _lib.dptr(self.seg_polys); h.seg_len = _lib.dptr(self.seg_len)
        h.n_crit = len(self.crit); h.crit_xy = _lib.dptr(self.crit)
        h.n_guide = len(self.guide); h.guide = _lib.dptr(self.guide)
        h.default_search_length = self.default_search_length
        for k in range(4):
            h.body_ext[k] = self.body_ext[k]
        h.n_aux = len(self.aux); h.aux_ext = _lib.dptr(self.aux)
        return h


def make_record(env, car, heuristic=None):
    """EnvRecord from the mirror objects (duck-typed: any object exposing the same
    attributes works)."""
    kw = {}
    if heuristic is not None:
        kw = dict(seg_xy=heuristic.seg_xy, seg_polys=heuristic.seg_polys, seg_len=heuristic.search_lengths,
                  crit_xy=heuristic.crit_xy, guide=heuristic.guided_path,
                  default_search_length=heuristic.default_search_length)
    return EnvRecord(env.obstacle_quads(), env.field_ring(), car.body_ext, car.aux_exts, **kw)


def pack_structs(records):
    """ctypes array of HlEnvHost pointing at the records' host buffers (reusable across uploads)."""
    return (_lib.HlEnvHost * len(records))(*[r.as_struct() for r in records])


class EnvBatch:
    """Device-resident environments.  ``records`` are EnvRecord objects."""

    def __init__(self, records, device=None, structs=None):
        self.records = list(records)
        self.ctx = _lib.get_ctx(device)         # device=None: the CALLING thread's current device (thread-local in CUDA)
        self.device = int(_lib.load_library().hl_ctx_device(self.ctx))
        arr = structs if structs is not None else pack_structs(self.records)
        h = C.c_void_p()
        _lib.check(_lib.load_library().hl_env_upload(self.ctx, arr, len(self.records), C.byref(h)),
                   "hl_env_upload")
        self.handle = h

    def __len__(self):
        return len(self.records)

    def close(self):
        if getattr(self, "handle", None):
            _lib.load_library().hl_env_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
