"""ctypes binding of libspeech_b200.so (include/speech_recog.h, include/sr_synth.h).

Plumbing for tests/ and bench.py only -- the product is the C-ABI library itself. Nothing here
computes: every call forwards to the CUDA kernels and raises if the library or a GPU is missing
(there is no CPU fallback). Struct layouts mirror the reference's VAD.H:10-22 / MFCC.H:18-25.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PKG_ROOT = os.path.normpath(os.path.join(_HERE, "..", ".."))
LIB_PATH = os.path.join(PKG_ROOT, "lib", "libspeech_b200.so")

FRAME_LEN, FRAME_MOV, MFCC_NUM, VV_FRM_MAX = 160, 80, 12, 119
FTR_BYTES = 2860
SEG_NULL = 0xFFFFFFFF
DIS_ERR = 0xFFFFFFFF
SAVE_MASK = 12345
DTW_CHECK_SIGN, DTW_BAND = 1, 2

ATAP_DTYPE = np.dtype([("mid_val", "<u4"), ("n_thl", "<u2"), ("z_thl", "<u2"), ("s_thl", "<u4")])
FTR_DTYPE = np.dtype([("save_sign", "<u2"), ("frm_num", "<u2"), ("mfcc_dat", "<i2", (VV_FRM_MAX * MFCC_NUM,))])
assert ATAP_DTYPE.itemsize == 12 and FTR_DTYPE.itemsize == FTR_BYTES


class RecogOut(C.Structure):
    _fields_ = [("atap", C.c_void_p), ("seg_off", C.c_void_p), ("ftr", C.c_void_p), ("score", C.c_void_p),
                ("best_idx", C.c_void_p), ("best_dis", C.c_void_p), ("cmd", C.c_void_p), ("status", C.c_void_p)]


class StreamEvent(C.Structure):
    _fields_ = [(k, C.c_uint32) for k in ("stream", "segment", "start", "end", "status", "frm_num", "best_idx", "best_dis", "cmd")]


class ValidTag(C.Structure):
    _fields_ = [("start", C.c_void_p), ("end", C.c_void_p)]


_lib = None


def lib():
    """The loaded shared library (raises if it has not been built: run `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libspeech_b200.so is not built (%s); run __graft_entry__.build()" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
        L.sr_create.argtypes = [i32, C.POINTER(vp)]
        L.sr_destroy.argtypes = [vp]
        L.sr_set_stream.argtypes = [vp, vp]
        L.sr_use_own_stream.argtypes = [vp]
        L.sr_sync.argtypes = [vp]
        L.sr_last_error.argtypes = [vp]
        L.sr_last_error.restype = C.c_char_p
        L.sr_host_alloc.argtypes = [C.c_size_t]
        L.sr_host_alloc.restype = vp
        L.sr_host_free.argtypes = [vp]
        L.sr_launch_count.argtypes = [vp]
        L.sr_launch_count.restype = u64
        L.sr_timing_enable.argtypes = [vp, u32]
        L.sr_timing_collect.argtypes = [vp, vp, vp, u32, vp]
        L.sr_debug_sqrt_mismatches.argtypes = [vp, u32, u32, vp]
        L.sr_set_transport.argtypes = [vp, C.c_int]
        L.sr_transport_stats.argtypes = [vp, vp, vp, vp]
        L.sr_debug_pack12_host.argtypes = [C.c_int, vp, u64, vp]
        L.sr_debug_pack12_host.restype = u32
        L.sr_debug_unpack12.argtypes = [vp, vp, u64, vp]
        L.sr_enrol_batch.argtypes = [vp, vp, u32, u32, u32, vp, u32, vp]
        L.sr_get_mdl_batch.argtypes = [vp, vp, vp, u32, vp, vp]
        L.sr_streams_create.argtypes = [vp, u32, u32, u32, C.POINTER(vp)]
        L.sr_streams_destroy.argtypes = [vp]
        L.sr_streams_reset.argtypes = [vp]
        L.sr_streams_push.argtypes = [vp, vp, u32, u32, vp, u32, vp]
        L.sr_streams_segments.argtypes = [vp, vp, vp]
        L.sr_streams_push_ragged.argtypes = [vp, vp, u32, vp, vp, u32, vp]
        L.sr_streams_fetch.argtypes = [vp, vp, u32, vp]
        L.sr_streams_pending.argtypes = [vp]
        L.sr_streams_pending.restype = u32
        L.sr_stream_group_create.argtypes = [C.POINTER(vp), u32, u32, u32, u32, C.POINTER(vp)]
        L.sr_stream_group_destroy.argtypes = [vp]
        L.sr_stream_group_reset.argtypes = [vp]
        L.sr_stream_group_push.argtypes = [vp, vp, u32, u32, vp, u32, vp]
        L.sr_stream_group_push_ragged.argtypes = [vp, vp, u32, vp, vp, u32, vp]
        L.sr_stream_group_segments.argtypes = [vp, vp, vp]
        L.sr_host_alloc_dev.argtypes = [i32, C.c_size_t]
        L.sr_host_alloc_dev.restype = vp
        L.sr_bind_thread_to_device.argtypes = [i32]
        L.sr_device_numa_node.argtypes = [i32]
        L.sr_host_numa_node.argtypes = [vp]
        L.sr_comm_unique_id.argtypes = [vp]
        L.sr_comm_create.argtypes = [vp, i32, i32, vp]
        L.sr_comm_destroy.argtypes = [vp]
        L.sr_comm_wait.argtypes = [vp]
        L.sr_allgather_dev.argtypes = [vp, vp, vp, C.c_size_t]
        L.sr_recognise_batch_dev_allgather.argtypes = [vp, vp, u32, u32, u32, C.POINTER(RecogOut), vp, vp]
        L.sr_set_dtw_variant.argtypes = [vp, i32]
        L.sr_set_geometry.argtypes = [vp, i32]
        L.sr_get_geometry.argtypes = [vp]
        L.sr_set_labels.argtypes = [vp, vp, u32, u32]
        L.sr_label.argtypes = [vp, u32]
        L.sr_label.restype = vp
        L.sr_set_bank.argtypes = [vp, vp, u32, u32]
        L.sr_set_bank_dev.argtypes = [vp, vp, u32, u32]
        for name in ("sr_noise_atap_batch", "sr_noise_atap_batch_dev"):
            getattr(L, name).argtypes = [vp, vp, u32, u32, u32, vp]
        for name in ("sr_vad_batch", "sr_vad_batch_dev"):
            getattr(L, name).argtypes = [vp, vp, u32, u32, u32, vp, vp]
        for name in ("sr_mfcc_batch", "sr_mfcc_batch_dev"):
            getattr(L, name).argtypes = [vp, vp, u32, u32, vp, u32, vp, vp]
        for name in ("sr_dtw_batch", "sr_dtw_batch_dev"):
            getattr(L, name).argtypes = [vp, vp, u32, u32, i32, vp, vp, vp]
        for name in ("sr_recognise_batch", "sr_recognise_batch_dev"):
            getattr(L, name).argtypes = [vp, vp, u32, u32, u32, C.POINTER(RecogOut)]
        L.sr_recognise_batch_multi.argtypes = [C.POINTER(vp), u32, vp, u32, u32, u32, C.POINTER(RecogOut)]
        L.sr_fft_mag_batch.argtypes = [vp, vp, u32, u32, vp]
        L.sr_fft_raw_batch.argtypes = [vp, vp, u32, vp]
        L.sr_get_dis_batch.argtypes = [vp, vp, vp, u32, vp]
        L.sr_dtw_limit_batch.argtypes = [vp, vp, vp, vp, vp, u32, vp]
        L.dtw_limit.argtypes = [C.c_uint16, C.c_uint16]
        L.dtw_limit.restype = C.c_uint8
        L.sr_synth_pcm_host.argtypes = [vp, u32, u32, u64, u32]
        L.sr_synth_pcm_dev.argtypes = [vp, u32, u32, u64, u32, vp]
        L.sr_synth_ftr_host.argtypes = [vp, u32, u32, u64, u32, u32]
        L.sr_wav_to_adc12.argtypes = [vp, C.c_size_t, vp, C.c_size_t, vp]
        L.sr_wav_to_adc12.restype = C.c_long
        L.noise_atap.argtypes = [vp, C.c_uint16, vp]
        L.noise_atap.restype = None
        L.VAD.argtypes = [vp, C.c_uint16, vp, vp]
        L.VAD.restype = None
        L.get_mfcc.argtypes = [vp, vp, vp]
        L.get_mfcc.restype = None
        L.dtw.argtypes = [vp, vp]
        L.dtw.restype = u32
        L.fft.argtypes = [vp, C.c_uint16]
        L.fft.restype = C.POINTER(C.c_uint32)
        L.get_dis.argtypes = [vp, vp]
        L.get_dis.restype = u32
        _lib = L
    return _lib


def _p(a):
    """void* of a numpy array (must be C-contiguous) or an int device pointer / None."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(C.c_void_p)
    return C.c_void_p(int(a))


class SrError(RuntimeError):
    pass


class Handle:
    """RAII wrapper of sr_handle. `device` is the CUDA ordinal."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        rc = lib().sr_create(int(device), C.byref(self._h))
        if rc != 0:
            raise SrError("sr_create failed (%d): %s" % (rc, lib().sr_last_error(None).decode()))
        self.n_slot = 0

    def close(self):
        if self._h:
            lib().sr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SrError("libspeech_b200 call failed (%d): %s" % (rc, lib().sr_last_error(self._h).decode()))

    # -- plumbing
    def set_stream(self, stream_ptr):
        self._ck(lib().sr_set_stream(self._h, _p(stream_ptr)))

    def use_own_stream(self):
        self._ck(lib().sr_use_own_stream(self._h))

    def sync(self):
        self._ck(lib().sr_sync(self._h))

    def launch_count(self):
        return int(lib().sr_launch_count(self._h))

    def set_transport(self, mode):
        """packed PCM transport of sr_recognise_batch: 0 off, 1 on, -1 automatic"""
        self._ck(lib().sr_set_transport(self._h, int(mode)))

    def transport_stats(self):
        """(packed chunks, plain chunks, bytes copied host -> device) of the last recognise() call on host buffers"""
        a, b, c = C.c_uint32(0), C.c_uint32(0), C.c_uint64(0)
        lib().sr_transport_stats(self._h, C.byref(a), C.byref(b), C.byref(c))
        return int(a.value), int(b.value), int(c.value)

    def unpack12(self, packed, n):
        """device expander of the packed transport alone (test hook): n samples from n/2*3 bytes"""
        packed = np.ascontiguousarray(packed, np.uint8)
        out = np.empty(n, np.uint16)
        self._ck(lib().sr_debug_unpack12(self._h, _p(packed.ctypes.data), n, _p(out.ctypes.data)))
        return out

    def timing_enable(self, max_records):
        self._ck(lib().sr_timing_enable(self._h, max_records))
        self._timing_cap = max_records

    def timing_collect(self):
        """[(tag, ms), ...] for every kernel launched since the last collect (synchronises the stream)"""
        cap = getattr(self, "_timing_cap", 0)
        tags, ms, n = np.zeros(cap, np.uint32), np.zeros(cap, np.float32), C.c_uint32(0)
        self._ck(lib().sr_timing_collect(self._h, _p(tags), _p(ms), cap, C.byref(n)))
        return list(zip(tags[: n.value].tolist(), ms[: n.value].tolist()))

    def set_bank(self, bank, n_slot, slot_stride):
        self._ck(lib().sr_set_bank(self._h, _p(bank), n_slot, slot_stride))
        self.n_slot = n_slot

    def set_bank_dev(self, bank_ptr, n_slot, slot_stride):
        self._ck(lib().sr_set_bank_dev(self._h, _p(bank_ptr), n_slot, slot_stride))
        self.n_slot = n_slot

    # -- host-buffer batched entry points (numpy in / numpy out)
    def noise_atap(self, pcm, n_len, atap=None):
        B, U = pcm.shape
        if atap is None:
            atap = np.zeros(B, ATAP_DTYPE)
        self._ck(lib().sr_noise_atap_batch(self._h, _p(pcm), U, B, n_len, _p(atap)))
        return atap

    def vad(self, pcm, atap, buf_len=None):
        B, U = pcm.shape
        seg = np.zeros((B, 3, 2), np.uint32)
        self._ck(lib().sr_vad_batch(self._h, _p(pcm), U, B, U if buf_len is None else buf_len, _p(atap), _p(seg)))
        return seg

    def mfcc(self, pcm, seg, atap, ftr=None):
        B, U = pcm.shape
        seg = np.ascontiguousarray(seg, np.uint32).reshape(B, -1)
        if ftr is None:
            ftr = np.zeros(B, FTR_DTYPE)
        self._ck(lib().sr_mfcc_batch(self._h, _p(pcm), U, B, _p(seg), seg.shape[1], _p(atap), _p(ftr)))
        return ftr

    def dtw(self, ftr_in, flags=0, band_r=0, want_score=True, want_best=True):
        B = ftr_in.shape[0]
        score = np.zeros((B, self.n_slot), np.uint32) if want_score else None
        bi = np.zeros(B, np.uint32) if want_best else None
        bd = np.zeros(B, np.uint32) if want_best else None
        self._ck(lib().sr_dtw_batch(self._h, _p(ftr_in), B, flags, band_r, _p(score), _p(bi), _p(bd)))
        return score, bi, bd

    def recognise(self, pcm, n_len=2400, want=("atap", "seg_off", "ftr", "score", "best_idx", "best_dis", "cmd", "status")):
        B, U = pcm.shape
        out = {}
        if "atap" in want:
            out["atap"] = np.zeros(B, ATAP_DTYPE)
        if "seg_off" in want:
            out["seg_off"] = np.zeros((B, 3, 2), np.uint32)
        if "ftr" in want:
            out["ftr"] = np.zeros(B, FTR_DTYPE)
        if "score" in want:
            out["score"] = np.zeros((B, self.n_slot), np.uint32)
        for k in ("best_idx", "best_dis", "cmd"):
            if k in want:
                out[k] = np.zeros(B, np.uint32)
        if "status" in want:
            out["status"] = np.zeros(B, np.uint8)
        ro = RecogOut(*[(_p(out[k]) if k in out else None) for k in
                        ("atap", "seg_off", "ftr", "score", "best_idx", "best_dis", "cmd", "status")])
        self._ck(lib().sr_recognise_batch(self._h, _p(pcm), U, B, n_len, C.byref(ro)))
        return out

    def enrol(self, pcm, n_len=2400, slot_stride=4096):
        B, U = pcm.shape
        bank = np.zeros((B, slot_stride), np.uint8)
        status = np.zeros(B, np.uint8)
        self._ck(lib().sr_enrol_batch(self._h, _p(pcm), U, B, n_len, _p(bank), slot_stride, _p(status)))
        return bank, status

    def get_mdl(self, in1, in2, mdl=None):
        n = in1.shape[0]
        if mdl is None:
            mdl = np.zeros(n, FTR_DTYPE)
        dis = np.zeros(n, np.uint32)
        self._ck(lib().sr_get_mdl_batch(self._h, _p(in1), _p(in2), n, _p(mdl), _p(dis)))
        return mdl, dis

    def fft_mag(self, frames):
        n, length = frames.shape
        mag = np.zeros((n, 512), np.uint32)
        self._ck(lib().sr_fft_mag_batch(self._h, _p(frames), length, n, _p(mag)))
        return mag

    def fft_raw(self, packed):
        n = packed.shape[0]
        out = np.zeros((n, 1024), np.uint32)
        self._ck(lib().sr_fft_raw_batch(self._h, _p(packed), n, _p(out)))
        return out

    def get_dis(self, a, b):
        n = a.shape[0]
        out = np.zeros(n, np.uint32)
        self._ck(lib().sr_get_dis_batch(self._h, _p(a), _p(b), n, _p(out)))
        return out

    # -- device-pointer entry points (ints / torch data_ptr()), asynchronous on the handle's stream
    def noise_atap_dev(self, pcm_ptr, U, B, n_len, atap_ptr):
        self._ck(lib().sr_noise_atap_batch_dev(self._h, _p(pcm_ptr), U, B, n_len, _p(atap_ptr)))

    def vad_dev(self, pcm_ptr, U, B, buf_len, atap_ptr, seg_ptr):
        self._ck(lib().sr_vad_batch_dev(self._h, _p(pcm_ptr), U, B, buf_len, _p(atap_ptr), _p(seg_ptr)))

    def mfcc_dev(self, pcm_ptr, U, B, seg_ptr, seg_stride, atap_ptr, ftr_ptr):
        self._ck(lib().sr_mfcc_batch_dev(self._h, _p(pcm_ptr), U, B, _p(seg_ptr), seg_stride, _p(atap_ptr), _p(ftr_ptr)))

    def dtw_dev(self, ftr_ptr, B, flags, band_r, score_ptr, bi_ptr, bd_ptr):
        self._ck(lib().sr_dtw_batch_dev(self._h, _p(ftr_ptr), B, flags, band_r, _p(score_ptr), _p(bi_ptr), _p(bd_ptr)))

    # -- the exchange step (NCCL behind the C-ABI)
    def comm_create(self, rank, world, id_bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(id_bytes))
        self._ck(lib().sr_comm_create(self._h, rank, world, buf))

    def comm_wait(self):
        self._ck(lib().sr_comm_wait(self._h))

    def allgather_dev(self, send_ptr, recv_ptr, nbytes):
        self._ck(lib().sr_allgather_dev(self._h, _p(send_ptr), _p(recv_ptr), nbytes))

    def recognise_dev_allgather(self, pcm_ptr, U, B, n_len, gathered_score=None, gathered_best=None, **ptrs):
        ro = RecogOut(*[_p(ptrs.get(k)) for k in
                        ("atap", "seg_off", "ftr", "score", "best_idx", "best_dis", "cmd", "status")])
        self._ck(lib().sr_recognise_batch_dev_allgather(self._h, _p(pcm_ptr), U, B, n_len, C.byref(ro),
                                                         _p(gathered_score), _p(gathered_best)))

    def set_dtw_variant(self, v):
        """greedy dtw kernel: 0 static lane = pair, 1 dynamic pair scheduling, -1 library default"""
        self._ck(lib().sr_set_dtw_variant(self._h, int(v)))

    def set_geometry(self, geom):
        """0 = reference 160/80/1024, 1 = GEOM_B 200/80/256 (extension, parity unpinned)"""
        self._ck(lib().sr_set_geometry(self._h, int(geom)))

    def label(self, cmd):
        """bytes of the command label (commstr, main.c:25-31) or None"""
        p = lib().sr_label(self._h, int(cmd))
        return C.string_at(p) if p else None

    def set_labels(self, labels, stride):
        raw = b"".join(bytes(l)[:stride - 1].ljust(stride, b"\0") for l in labels)
        self._ck(lib().sr_set_labels(self._h, raw, len(labels), stride))

    def recognise_dev(self, pcm_ptr, U, B, n_len, **ptrs):
        ro = RecogOut(*[_p(ptrs.get(k)) for k in
                        ("atap", "seg_off", "ftr", "score", "best_idx", "best_dis", "cmd", "status")])
        self._ck(lib().sr_recognise_batch_dev(self._h, _p(pcm_ptr), U, B, n_len, C.byref(ro)))


def comm_unique_id():
    """128 bytes identifying a new communicator (call on one rank, hand to the others)"""
    buf = (C.c_uint8 * 128)()
    rc = lib().sr_comm_unique_id(buf)
    if rc != 0:
        raise SrError("sr_comm_unique_id failed (%d): %s" % (rc, lib().sr_last_error(None).decode()))
    return bytes(buf)


def recognise_multi(handles, pcm, n_len=2400, want=("best_idx", "best_dis", "cmd", "status", "score", "seg_off")):
    """sr_recognise_batch_multi: one host call over several handles (one per GPU); numpy in/out"""
    B, U = pcm.shape
    T = handles[0].n_slot
    out = {}
    if "seg_off" in want:
        out["seg_off"] = np.zeros((B, 3, 2), np.uint32)
    if "ftr" in want:
        out["ftr"] = np.zeros(B, FTR_DTYPE)
    if "score" in want:
        out["score"] = np.zeros((B, T), np.uint32)
    for k in ("best_idx", "best_dis", "cmd"):
        if k in want:
            out[k] = np.zeros(B, np.uint32)
    if "status" in want:
        out["status"] = np.zeros(B, np.uint8)
    ro = RecogOut(*[(_p(out[k]) if k in out else None) for k in
                    ("atap", "seg_off", "ftr", "score", "best_idx", "best_dis", "cmd", "status")])
    arr = (C.c_void_p * len(handles))(*[h._h for h in handles])
    rc = lib().sr_recognise_batch_multi(arr, len(handles), _p(pcm), U, B, n_len, C.byref(ro))
    if rc != 0:
        raise SrError("sr_recognise_batch_multi failed (%d): %s" % (rc, lib().sr_last_error(None).decode()))
    return out


class StreamPool:
    """sr_stream_pool / sr_stream_group wrapper: chunked capture of S streams, in lock step or ragged
    (include/speech_recog.h, streaming section). `handle` may be a list of handles: the streams are then sharded over
    them (one GPU each)."""

    def __init__(self, handle, n_streams, max_samples, n_len=2400):
        self.S, self.L = n_streams, max_samples
        self._p = C.c_void_p()
        self.group = isinstance(handle, (list, tuple))
        self.h = handle[0] if self.group else handle
        if self.group:
            arr = (C.c_void_p * len(handle))(*[h._h for h in handle])
            rc = lib().sr_stream_group_create(arr, len(handle), n_streams, max_samples, n_len, C.byref(self._p))
            if rc != 0:
                raise SrError("sr_stream_group_create failed (%d): %s" % (rc, lib().sr_last_error(None).decode()))
        else:
            handle._ck(lib().sr_streams_create(handle._h, n_streams, max_samples, n_len, C.byref(self._p)))
        self._ev = (StreamEvent * (3 * n_streams))()

    def _ck(self, rc):
        if rc != 0:
            raise SrError("streaming call failed (%d): %s" % (rc, (lib().sr_last_error(None) or b"").decode()))

    def close(self):
        if self._p:
            (lib().sr_stream_group_destroy if self.group else lib().sr_streams_destroy)(self._p)
            self._p = C.c_void_p()

    def reset(self):
        self._ck((lib().sr_stream_group_reset if self.group else lib().sr_streams_reset)(self._p))

    def _events(self, n):
        return [{k: getattr(self._ev[i], k) for k, _ in StreamEvent._fields_} for i in range(n)]

    def push(self, chunk, chunk_len=None, stride=None, max_events=None):
        """chunk: numpy [S, chunk_len] u16 (or a raw host pointer with chunk_len/stride). Returns list of event dicts."""
        if isinstance(chunk, np.ndarray):
            chunk_len, stride, ptr = chunk.shape[1], chunk.strides[0] // 2, chunk.ctypes.data_as(C.c_void_p)
        else:
            ptr = C.c_void_p(int(chunk))
        n = C.c_uint32(0)
        f = lib().sr_stream_group_push if self.group else lib().sr_streams_push
        self._ck(f(self._p, ptr, chunk_len, stride, self._ev, 3 * self.S if max_events is None else max_events, C.byref(n)))
        return self._events(n.value)

    def push_raw(self, ptr, chunk_len, stride, max_events=None):
        """lock-step push from a raw host pointer; returns only the NUMBER of events (they stay in self._ev): what a
        latency measurement should time -- no Python object is built per event"""
        n = C.c_uint32(0)
        f = lib().sr_stream_group_push if self.group else lib().sr_streams_push
        self._ck(f(self._p, C.c_void_p(int(ptr)), chunk_len, stride, self._ev, 3 * self.S if max_events is None else max_events, C.byref(n)))
        return n.value

    def push_ragged(self, chunk, lens, stride=None, max_events=None):
        """chunk: numpy [S, >= max(lens)] u16 (or raw pointer + stride); lens: [S] samples for each stream"""
        lens = np.ascontiguousarray(lens, np.uint32)
        if isinstance(chunk, np.ndarray):
            stride, ptr = chunk.strides[0] // 2, chunk.ctypes.data_as(C.c_void_p)
        else:
            ptr = C.c_void_p(int(chunk))
        n = C.c_uint32(0)
        f = lib().sr_stream_group_push_ragged if self.group else lib().sr_streams_push_ragged
        self._ck(f(self._p, ptr, stride, _p(lens), self._ev, 3 * self.S if max_events is None else max_events, C.byref(n)))
        return self._events(n.value)

    def fetch(self, max_events=None):
        assert not self.group
        n = C.c_uint32(0)
        self._ck(lib().sr_streams_fetch(self._p, self._ev, 3 * self.S if max_events is None else max_events, C.byref(n)))
        return self._events(n.value)

    def pending(self):
        return int(lib().sr_streams_pending(self._p))

    def segments(self):
        seg = np.zeros((self.S, 3, 2), np.uint32)
        atap = np.zeros(self.S, ATAP_DTYPE)
        f = lib().sr_stream_group_segments if self.group else lib().sr_streams_segments
        self._ck(f(self._p, _p(seg), _p(atap)))
        return seg, atap


def host_alloc_dev(device, nbytes):
    """pinned host memory on `device`'s NUMA node as a numpy uint8 array (keeps the allocation alive through .base)"""
    p = lib().sr_host_alloc_dev(int(device), nbytes)
    if not p:
        raise SrError("sr_host_alloc_dev(%d, %d) failed" % (device, nbytes))
    buf = (C.c_uint8 * nbytes).from_address(p)
    arr = np.frombuffer(buf, np.uint8)
    return arr, p


def host_free(p):
    lib().sr_host_free(C.c_void_p(p))


# ---- synthetic workload (include/sr_synth.h) -------------------------------------------------------
def synth_pcm_host(B, U, seed_base, nwords=1):
    pcm = np.zeros((B, U), np.uint16)
    rc = lib().sr_synth_pcm_host(_p(pcm), U, B, seed_base, nwords)
    if rc != 0:
        raise SrError("sr_synth_pcm_host failed")
    return pcm


def synth_pcm_dev(pcm_ptr, B, U, seed_base, nwords=1, stream_ptr=None):
    rc = lib().sr_synth_pcm_dev(_p(pcm_ptr), U, B, seed_base, nwords, _p(stream_ptr))
    if rc != 0:
        raise SrError("sr_synth_pcm_dev failed")


def synth_ftr_host(B, seed_base, fmin=50, fmax=100, stride=FTR_BYTES):
    buf = np.zeros((B, stride), np.uint8)
    rc = lib().sr_synth_ftr_host(_p(buf), stride, B, seed_base, fmin, fmax)
    if rc != 0:
        raise SrError("sr_synth_ftr_host failed")
    return buf


def wav_to_adc12(wav_bytes, max_samples=1 << 24):
    """(samples u16[n], sample_rate) from the bytes of a RIFF/WAVE PCM file (include/sr_synth.h)"""
    buf = np.frombuffer(wav_bytes, np.uint8).copy()
    out = np.zeros(min(max_samples, max(len(buf), 1)), np.uint16)
    rate = C.c_uint32(0)
    n = lib().sr_wav_to_adc12(_p(buf), len(buf), _p(out), len(out), C.byref(rate))
    if n < 0:
        raise SrError("not a supported PCM WAV file")
    return out[:n].copy(), rate.value


def make_bank(ftr, slot_stride=4096, valid=None):
    """Pack feature structs into flash-layout slots (Src/BSP/Flash.H:11-20, Flash.C:41-63): save_sign =
    save_mask, frm_num, rows; the rest of the slot stays erased-flash 0xFF."""
    T = ftr.shape[0]
    bank = np.full((T, slot_stride), 0xFF, np.uint8)
    raw = ftr.view(np.uint8).reshape(T, FTR_BYTES)
    for t in range(T):
        n = int(ftr["frm_num"][t])
        bank[t, 2:4 + 24 * n] = raw[t, 2:4 + 24 * n]
        sign = SAVE_MASK if (valid is None or valid[t]) else 0xFFFF
        bank[t, 0] = sign & 0xFF
        bank[t, 1] = sign >> 8
    return bank


def pack12_host(x, variant=-1):
    """host packer of the packed transport alone (test hook, no GPU): returns (packed bytes, OR of all samples) or None if
    the variant is not available on this CPU"""
    x = np.ascontiguousarray(x, np.uint16)
    dst = np.zeros(x.size // 2 * 3, np.uint8)
    o = lib().sr_debug_pack12_host(int(variant), _p(x.ctypes.data), x.size, _p(dst.ctypes.data))
    return None if o == 0xFFFFFFFF else (dst, int(o))
