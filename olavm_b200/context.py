"""Context = one GPU's prover state (streams, twiddles).  Mirrors the reference's one-time `init_gpu()`
(plonky2/field/src/cfft/ntt/mod.rs:55-101, called from OlaStark::default(), circuits/src/stark/ola_stark.rs:47)."""
import ctypes

import numpy as np

from . import _lib


class Context:
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self._lib.ola_gpu_init(int(device), ctypes.byref(h))
        if rc != 0:
            raise _lib.OlaError(rc, f"ola_gpu_init(device={device}) failed: a B200 (sm_100) GPU is required; no CPU fallback")
        self.handle = h
        self.device = device

    # ---- plumbing
    def check(self, rc):
        if rc < 0:
            msg = self._lib.ola_gpu_last_error(self.handle)
            raise _lib.OlaError(rc, msg.decode() if msg else "")
        return rc

    # ---- C::Hasher of the GenericConfig (plonk/config.rs:115-122 Poseidon, :153-161 Blake3)
    @property
    def hasher(self):
        return int(self._lib.ola_get_hasher(self.handle))

    @hasher.setter
    def hasher(self, hasher_id):
        self.check(self._lib.ola_set_hasher(self.handle, int(hasher_id)))

    def sync(self):
        self.check(self._lib.ola_gpu_sync(self.handle))

    @property
    def kernel_launches(self):
        return int(self._lib.ola_gpu_kernel_launches(self.handle))

    @property
    def stream_ptr(self):
        return int(self._lib.ola_gpu_stream(self.handle) or 0)

    def profile_begin(self):
        self.check(self._lib.ola_profile_begin(self.handle))

    def profile_end(self):
        """-> {kernel name: {"ms": total device time, "launches": n}} since profile_begin()."""
        import json

        buf = ctypes.create_string_buffer(1 << 16)
        self.check(self._lib.ola_profile_end(self.handle, buf, len(buf)))
        return json.loads(buf.value.decode())

    def close(self):
        if getattr(self, "handle", None):
            self._lib.ola_gpu_destroy(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- raw device buffers
    def alloc(self, n_u64):
        p = ctypes.c_void_p()
        self.check(self._lib.ola_dev_alloc(self.handle, int(n_u64), ctypes.byref(p)))
        return p

    def free(self, p):
        self.check(self._lib.ola_dev_free(self.handle, p))

    def upload(self, host, dev=None):
        host = np.ascontiguousarray(host, dtype=np.uint64)
        if dev is None:
            dev = self.alloc(host.size)
        self.check(self._lib.ola_dev_upload(self.handle, dev, _lib.hptr(host), host.size))
        return dev

    def download(self, dev, shape):
        out = np.empty(shape, dtype=np.uint64)
        self.check(self._lib.ola_dev_download(self.handle, _lib.hptr(out), dev, out.size))
        return out
