"""Host-side mirror of the reference interface for the marcher path.

`TraceRaysCu` mirrors the reference's boundary class of the same name (src/cuda_volume_raytracer.h:61-115:
ctor(bounds, diff planes, translucency_cropped), trace_rays_cu(...), public _output_sizes);
`RaytraceScene` mirrors the scene class above it (src/image_util.h:132-195, image_util.cpp:501-772: ctor(bounds,
ior, translucency), trace_rays(...) in API coordinates with the normalise step and the +-0x10000 shifts), with
prep and normalisation running on the GPU ("next" rows f1/f2).  Both are thin: every number is produced by
libvrt_b200.so through the C ABI.  numpy arrays in, numpy arrays out; `*_device` variants take torch CUDA tensors.
"""
import ctypes as C

import numpy as np

from . import _lib as L

_DT = {np.dtype(np.float32): L.VRT_F32, np.dtype(np.int16): L.VRT_I16, np.dtype(np.uint32): L.VRT_U32}


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Options:
    """Mirror of the reference's Options (types.h:83-91).  _loglevel/_minimum_gpu/_max_cpu are accepted for
    source compatibility; there is no CPU path, so _minimum_gpu and _max_cpu have no effect."""

    def __init__(self, loglevel=0, minimum_gpu=0x80, write_instance=False):
        self._loglevel = loglevel
        self._minimum_gpu = minimum_gpu
        self._write_instance = write_instance
        self._max_cpu = 256


class TraceRaysCu:
    """B200 drop-in for TraceRaysCu<DiffType> (DiffType from the planes' dtype: float32 or int16)."""

    def __init__(self, output_sizes, diff, translucency_cropped, device=0, _handle=None, bricked=False, keep_i16=False, texture=False, paired=False):
        self._h = C.c_void_p()
        self._keepalive = None
        if _handle is not None:
            self._h = _handle
        else:
            bounds = np.asarray(output_sizes, dtype=np.uint64)
            dim = len(bounds)
            planes = [np.ascontiguousarray(d).reshape(-1) for d in diff]
            if len(planes) != dim:
                raise ValueError("need one gradient plane per axis")
            dt = _DT.get(planes[0].dtype)
            if dt not in (L.VRT_F32, L.VRT_I16):
                raise TypeError("diff planes must be float32 or int16")
            tr = np.ascontiguousarray(translucency_cropped, dtype=np.uint32).reshape(-1)
            nvox = int(np.prod(bounds.astype(object)))
            if any(p.size != nvox for p in planes) or tr.size != nvox:
                raise ValueError("imagesizes doesn't match")
            ptrs = (C.c_void_p * dim)(*[p.ctypes.data for p in planes])
            L.check(L.lib().vrt_scene_create(C.byref(self._h), device, dim, _p(bounds), dt, ptrs, _p(tr),
                                             (L.VRT_SCENE_LAYOUT_BRICK if bricked else 0) | (L.VRT_SCENE_KEEP_I16 if keep_i16 else 0) | (L.VRT_SCENE_LAYOUT_TEXTURE if texture else 0) | (L.VRT_SCENE_LAYOUT_PAIR if paired else 0)))
        self._read_info()

    # -- alternative constructors -------------------------------------------------------------------
    @classmethod
    def from_interleaved(cls, output_sizes, volume, translucency_cropped, device=0):
        bounds = np.asarray(output_sizes, dtype=np.uint64)
        volume = np.ascontiguousarray(volume)
        tr = np.ascontiguousarray(translucency_cropped, dtype=np.uint32).reshape(-1)
        h = C.c_void_p()
        L.check(L.lib().vrt_scene_create_interleaved(C.byref(h), device, len(bounds), _p(bounds), _DT[volume.dtype], _p(volume), _p(tr), 0))
        return cls(None, None, None, _handle=h)

    @classmethod
    def from_device(cls, output_sizes, volume_tensor, translucency_tensor=None, borrow=True):
        """volume_tensor: torch CUDA tensor [nvox, dim+1] float32|int16 (e.g. just filled by an NCCL broadcast)."""
        import torch
        bounds = np.asarray(output_sizes, dtype=np.uint64)
        dt = L.VRT_F32 if volume_tensor.dtype == torch.float32 else L.VRT_I16
        h = C.c_void_p()
        trp = C.c_void_p(translucency_tensor.data_ptr()) if translucency_tensor is not None else None
        L.check(L.lib().vrt_scene_create_device(C.byref(h), volume_tensor.device.index or 0, len(bounds), _p(bounds), dt,
                                                C.c_void_p(volume_tensor.data_ptr()), trp, L.VRT_SCENE_BORROW if borrow else 0))
        obj = cls(None, None, None, _handle=h)
        if borrow:
            obj._keepalive = (volume_tensor, translucency_tensor)
        return obj

    @classmethod
    def from_ior(cls, bound_vec, ior, translucency, device=0, bricked=False, keep_i16=False, texture=False, paired=False):
        """GPU scene prep (f1).  ior: numpy float32|uint32 or a torch CUDA tensor of those; bounds UNCROPPED."""
        bounds = np.asarray(bound_vec, dtype=np.uint64)
        h = C.c_void_p()
        if isinstance(ior, np.ndarray):
            ior = np.ascontiguousarray(ior).reshape(-1)
            tr = np.ascontiguousarray(translucency, dtype=np.uint32).reshape(-1)
            if ior.size != int(np.prod(bounds.astype(object))) or tr.size != ior.size:
                raise VrtErrorCompat("imagesizes doesn't match")
            L.check(L.lib().vrt_scene_create_from_ior(C.byref(h), device, len(bounds), _p(bounds), _DT[ior.dtype], _p(ior), _p(tr), 0,
                                                      (L.VRT_SCENE_LAYOUT_BRICK if bricked else 0) | (L.VRT_SCENE_KEEP_I16 if keep_i16 else 0) | (L.VRT_SCENE_LAYOUT_TEXTURE if texture else 0) | (L.VRT_SCENE_LAYOUT_PAIR if paired else 0)))
        else:
            import torch
            dt = L.VRT_F32 if ior.dtype == torch.float32 else L.VRT_U32
            L.check(L.lib().vrt_scene_create_from_ior(C.byref(h), ior.device.index or 0, len(bounds), _p(bounds), dt,
                                                      C.c_void_p(ior.data_ptr()), C.c_void_p(translucency.data_ptr()), 1,
                                                      (L.VRT_SCENE_LAYOUT_BRICK if bricked else 0) | (L.VRT_SCENE_KEEP_I16 if keep_i16 else 0) | (L.VRT_SCENE_LAYOUT_TEXTURE if texture else 0) | (L.VRT_SCENE_LAYOUT_PAIR if paired else 0)))
        return cls(None, None, None, _handle=h)

    # -- plumbing -----------------------------------------------------------------------------------
    def _read_info(self):
        dev, dim, dt, nbytes = C.c_int(), C.c_int(), C.c_int(), C.c_uint64()
        bounds = np.zeros(3, dtype=np.uint64)
        dvol, dtr = C.c_void_p(), C.c_void_p()
        L.check(L.lib().vrt_scene_info(self._h, C.byref(dev), C.byref(dim), _p(bounds), C.byref(dt), C.byref(dvol), C.byref(dtr), C.byref(nbytes)))
        self.device, self.dim, self.diff_dtype = dev.value, dim.value, dt.value
        self._output_sizes = [int(b) for b in bounds[:dim.value]]      # public member in the reference (h:73)
        self.volume_ptr, self.translucency_ptr, self.volume_bytes = dvol.value, dtr.value, nbytes.value

    def storage_info(self):
        """(storage dtype, layout flags, bytes, device address) of the staged copy (vrt_scene_storage_info)."""
        dt, fl, nb, ptr = C.c_int(), C.c_uint(), C.c_uint64(), C.c_void_p()
        L.check(L.lib().vrt_scene_storage_info(self._h, C.byref(dt), C.byref(fl), C.byref(nb), C.byref(ptr)))
        return dt.value, fl.value, nb.value, ptr.value

    def replicate(self, devices):
        """In-process multi-GPU replication over NVLink peer copies (vrt_scene_replicate; reference: the per-device host
        upload loop cu:676-686).  Returns ([TraceRaysCu on devices[i]], seconds)."""
        devs = (C.c_int * len(devices))(*devices)
        outs = (C.c_void_p * len(devices))()
        secs = C.c_double(0.0)
        L.check(L.lib().vrt_scene_replicate(self._h, len(devices), devs, outs, C.byref(secs)))
        return [TraceRaysCu(None, None, None, _handle=C.c_void_p(h)) for h in outs], secs.value

    def download_volume(self):
        """Host copy of the staged interleaved volume [nvox, dim+1] and of the cropped translucency plane."""
        nvox = int(np.prod(self._output_sizes))
        vol = np.empty((nvox, self.dim + 1), dtype=np.float32 if self.diff_dtype == L.VRT_F32 else np.int16)
        tr = np.empty(nvox, dtype=np.uint32)
        L.check(L.lib().vrt_scene_download(self._h, _p(vol), _p(tr)))
        return vol, tr

    def export_device(self, volume_tensor, translucency_tensor=None, stream=None):
        """D2D copy of the staged volume into torch CUDA tensors (the buffers an NCCL broadcast then replicates)."""
        import torch
        st = torch.cuda.current_stream(volume_tensor.device) if stream is None else stream
        L.check(L.lib().vrt_scene_export_device(self._h, C.c_void_p(volume_tensor.data_ptr()),
                                                C.c_void_p(translucency_tensor.data_ptr()) if translucency_tensor is not None else None,
                                                C.c_void_p(st.cuda_stream)))

    def set_option(self, key, value):
        L.check(L.lib().vrt_scene_set_option(self._h, key, int(value)))

    def get_option(self, key):
        v = C.c_int64()
        L.check(L.lib().vrt_scene_get_option(self._h, key, C.byref(v)))
        return v.value

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            L.lib().vrt_scene_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- the path -----------------------------------------------------------------------------------
    def trace_rays_cu(self, start_position, start_direction, scale_vec, minimum_brightness, iterations,
                      trace_paths=False, opt=None, live_translucency=False, round_host=False):
        """Returns (end_position, end_direction, end_iteration, remaining_light, path).  Coordinates are
        cropped-volume 16.16 (the reference boundary's convention).  live_translucency=False reproduces the shipped
        behaviour, where the per-step attenuation is compiled out and minimum_brightness is ignored (cu:785)."""
        dim = self.dim
        pos = np.ascontiguousarray(start_position, dtype=np.uint32).reshape(-1)
        d = np.ascontiguousarray(start_direction).reshape(-1)
        if d.dtype not in (np.float32, np.int16):
            raise TypeError("directions must be float32 or int16")
        if pos.size % dim or pos.size != d.size:
            raise L.VrtError(L.VRT_ERR_INVALID, "raycounts doesn't match")
        n = pos.size // dim
        isc = np.ascontiguousarray(scale_vec, dtype=np.float32)
        if isc.size < dim:
            raise L.VrtError(L.VRT_ERR_INVALID, "scale_vec too short")
        epos = np.empty_like(pos); edir = np.empty_like(d)
        eit = np.empty(n, dtype=np.uint32); light = np.empty(n, dtype=np.uint32)
        path = np.empty(n * iterations * dim, dtype=np.uint32) if trace_paths else None
        flags = (L.VRT_TRACE_PATHS if trace_paths else 0) | (L.VRT_TRACE_LIVE_TRANSLUCENCY if live_translucency else 0) | \
                (L.VRT_TRACE_ROUND_HOST if round_host else 0)
        L.check(L.lib().vrt_trace(self._h, n, _p(pos), _p(d), _DT[d.dtype], _p(isc), int(minimum_brightness), int(iterations), flags,
                                  _p(epos), _p(edir), _p(eit), _p(light), _p(path)))
        if trace_paths:
            path = path.reshape(n, iterations, dim)
        return epos.reshape(n, dim), edir.reshape(n, dim), eit, light, path

    def trace_host_buffers(self, pos, d, isc, minimum_brightness, iterations, epos, edir, eit, light, flags=0):
        """Raw vrt_trace on caller-owned host arrays / pinned torch tensors (used by bench.py's e2e leg)."""
        def ptr(x):
            return C.c_void_p(x.ctypes.data) if isinstance(x, np.ndarray) else C.c_void_p(x.data_ptr())
        dt = L.VRT_I16 if (d.dtype == np.int16 if isinstance(d, np.ndarray) else d.element_size() == 2) else L.VRT_F32
        n = (pos.size if isinstance(pos, np.ndarray) else pos.numel()) // self.dim
        isc = np.ascontiguousarray(isc, dtype=np.float32)
        L.check(L.lib().vrt_trace(self._h, n, ptr(pos), ptr(d), dt, _p(isc), int(minimum_brightness), int(iterations), flags,
                                  ptr(epos), ptr(edir), ptr(eit), ptr(light), None))

    @staticmethod
    def cap_hit():
        """1 / 0 / -1: did any ray of this thread's last host-buffer trace end at the iteration cap (vrt_trace_cap_hit)."""
        return int(L.lib().vrt_trace_cap_hit())

    def trace_device(self, pos, d, scale_vec, minimum_brightness, iterations, epos=None, edir=None, eit=None, light=None,
                     path=None, live_translucency=False, stream=None, round_host=False):
        """torch CUDA tensors in/out, enqueued on `stream` (default: torch's current stream); does not synchronise."""
        import torch
        n = pos.numel() // self.dim
        epos = torch.empty_like(pos) if epos is None else epos
        edir = torch.empty_like(d) if edir is None else edir
        eit = torch.empty(n, dtype=torch.int32, device=pos.device) if eit is None else eit
        light = torch.empty(n, dtype=torch.int32, device=pos.device) if light is None else light
        isc = np.ascontiguousarray(scale_vec, dtype=np.float32)
        st = torch.cuda.current_stream(pos.device) if stream is None else stream
        flags = (L.VRT_TRACE_PATHS if path is not None else 0) | (L.VRT_TRACE_LIVE_TRANSLUCENCY if live_translucency else 0) | \
                (L.VRT_TRACE_ROUND_HOST if round_host else 0)
        dt = L.VRT_I16 if d.dtype == torch.int16 else L.VRT_F32
        L.check(L.lib().vrt_trace_device(self._h, n, C.c_void_p(pos.data_ptr()), C.c_void_p(d.data_ptr()), dt, _p(isc),
                                         int(minimum_brightness), int(iterations), flags, C.c_void_p(epos.data_ptr()),
                                         C.c_void_p(edir.data_ptr()), C.c_void_p(eit.data_ptr()), C.c_void_p(light.data_ptr()),
                                         C.c_void_p(path.data_ptr()) if path is not None else None, C.c_void_p(st.cuda_stream)))
        return epos, edir, eit, light

    def normalise_rays_device(self, pos, d, stream=None):
        import torch
        st = torch.cuda.current_stream(pos.device) if stream is None else stream
        bad = C.c_int64(0)
        dt = L.VRT_I16 if d.dtype == torch.int16 else L.VRT_F32
        L.check(L.lib().vrt_normalise_rays_device(self._h, pos.numel() // self.dim, C.c_void_p(pos.data_ptr()), C.c_void_p(d.data_ptr()),
                                                  dt, C.byref(bad), C.c_void_p(st.cuda_stream)))


class Comm:
    """One process per GPU: NCCL communicator owned by libvrt_b200.so (vrt_comm_*).  `exchange(bytes_or_None) -> bytes` hands
    rank 0's 128-byte unique id to every rank (e.g. torch.distributed.broadcast_object_list, MPI, a file)."""

    def __init__(self, device, rank, world, exchange):
        self._c = C.c_void_p()
        # libvrt_b200.so dlopens "libnccl.so.2".  In a process that also uses PyTorch, torch's bundled NCCL must be the copy that gets
        # mapped (the loader keeps one object per soname, and libtorch_cuda.so needs symbols a system-wide older NCCL may lack): make
        # sure torch is loaded first when it is installed.
        try:
            import torch  # noqa: F401
        except ImportError:
            pass
        uid = None
        if rank == 0:
            buf = C.create_string_buffer(128)
            L.check(L.lib().vrt_comm_unique_id(buf))
            uid = buf.raw
        uid = exchange(uid)
        self.rank, self.world, self.device = rank, world, device
        L.check(L.lib().vrt_comm_create(C.byref(self._c), device, rank, world, C.c_char_p(uid)))

    def broadcast_scene(self, scene, root=0):
        """Replicates root's scene to every rank with in-place ncclBroadcasts of the staged buffers (vrt_scene_broadcast).
        Returns (TraceRaysCu on this rank, seconds of the payload broadcasts on the device)."""
        out = C.c_void_p()
        secs = C.c_double(0.0)
        L.check(L.lib().vrt_scene_broadcast(self._c, root, scene._h if scene is not None else None, C.byref(out), C.byref(secs)))
        if scene is not None and self.rank == root:
            return scene, secs.value
        return TraceRaysCu(None, None, None, _handle=out), secs.value

    def close(self):
        if self._c:
            L.lib().vrt_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def VrtErrorCompat(msg):
    return L.VrtError(L.VRT_ERR_INVALID, msg)


class RaytraceScene:
    """Mirror of RaytraceScene<IorType, IorLogType, DiffType> (image_util.h:132-195): ior float32 -> float scene,
    ior uint32 (16.16) -> int16 scene.  Scene prep (f1) and ray normalisation (f2) run on the GPU."""

    def __init__(self, bound_vec, ior, translucency, opt=None, device=0):
        self._bound_vec = [int(b) for b in bound_vec]
        if len(self._bound_vec) == 0:
            raise L.VrtError(L.VRT_ERR_INVALID, "dimension is zero")
        self._calculation_object = TraceRaysCu.from_ior(self._bound_vec, ior, translucency, device=device)
        self.dim = len(self._bound_vec)
        self.float_scene = self._calculation_object.diff_dtype == L.VRT_F32

    def close(self):
        self._calculation_object.close()

    def trace_rays(self, start_position, start_direction, scale, minimum_brightness, iterations, trace_path=False,
                   normalize_length=True, opt=None, live_translucency=False):
        """API coordinates in, API coordinates out (image_util.cpp:645-772)."""
        import torch
        dev = torch.device("cuda", self._calculation_object.device)
        dim = self.dim
        pos = np.ascontiguousarray(start_position, dtype=np.uint32).reshape(-1)
        d = np.ascontiguousarray(start_direction, dtype=np.float32 if self.float_scene else np.int16).reshape(-1)
        if pos.size % dim or pos.size != d.size:
            raise L.VrtError(L.VRT_ERR_INVALID, "raycounts doesn't match, dimension is: %d raysizes are start_position: %d start_direction: %d"
                             % (dim, pos.size, d.size))
        n = pos.size // dim
        tpos = torch.from_numpy(pos.view(np.int32)).to(dev)
        tdir = torch.from_numpy(d).to(dev)
        self._calculation_object.normalise_rays_device(tpos, tdir)                     # image_util.cpp:675-719
        tpath = torch.empty(n * iterations * dim, dtype=torch.int32, device=dev) if trace_path else None
        epos, edir, eit, light = self._calculation_object.trace_device(tpos, tdir, scale, minimum_brightness, iterations,
                                                                        path=tpath, live_translucency=live_translucency)
        epos += 0x10000                                                                 # image_util.cpp:771
        out_path = None
        if trace_path:
            tpath += 0x10000                                                            # image_util.cpp:770
            out_path = tpath.cpu().numpy().view(np.uint32).reshape(n, iterations, dim)
        return (epos.cpu().numpy().view(np.uint32).reshape(n, dim), edir.cpu().numpy().reshape(n, dim),
                eit.cpu().numpy().view(np.uint32), light.cpu().numpy().view(np.uint32), out_path)
