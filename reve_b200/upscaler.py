"""Host-side mirror of the reference's upscale stage over the C ABI.

``Upscaler`` is one ``reve_ctx`` (one GPU); ``upscale_segment`` mirrors
``Video::upscale_segment`` (reference reve-shared/src/lib.rs:129-155): a directory of decoded
frames in, the same file names upscaled out, one ``"<in> -> <out> done"`` line per frame on the
progress stream (what reve-cli/src/main.rs:265-273 counts).  Every computation happens in
libreve_cuda.so; nothing here computes pixels on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib

REVE_E_BUSY = -7
FMT_RGB24, FMT_YUV420P10LE_BT601, FMT_YUV420P10LE_BT709 = 0, 1, 2
CTX_SHARED_DEVICE = 1
# test hooks (reve_ctx_options.debug_flags, include/reve_cuda.h)
DBG_NO_REVERSE, DBG_CTA_PAIRS, DBG_SWAP_PAIR_B, DBG_ALL_ROWS, DBG_ALIAS_ROWS, DBG_FAULT, DBG_EQUAL_SPLIT = 1, 2, 4, 8, 16, 32, 64
DBG_CONV0_IM2COL = 128


class ReveError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[reve status {status}] {msg}")
        self.status = status


def load_library():
    return _lib.load()


def library_path() -> str:
    return _lib.LIB_PATH


def _check(rc: int, ctx=None):
    if rc != 0:
        lib = _lib.load()
        msg = lib.reve_last_error(ctx).decode("utf-8", "replace")
        if not msg:
            msg = lib.reve_strerror(rc).decode()
        raise ReveError(rc, msg)


class Model:
    """Weights of realesr-animevideov3-x{2,3,4} (replaces ``-n <name>`` + ``models\\``)."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def load_ncnn(cls, param_path: str, bin_path: str) -> "Model":
        h = C.c_void_p()
        _check(_lib.load().reve_model_load_ncnn(param_path.encode(), bin_path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def random(cls, scale: int, seed: int) -> "Model":
        h = C.c_void_p()
        _check(_lib.load().reve_model_random(scale, seed, C.byref(h)))
        return cls(h)

    @classmethod
    def from_arrays(cls, scale: int, conv_w: Sequence[np.ndarray], conv_b: Sequence[np.ndarray],
                    prelu: Sequence[np.ndarray]) -> "Model":
        """18 OIHW conv weights, 18 biases, 17 PReLU slope vectors (fp32; copied by the library)."""
        if len(conv_w) != 18 or len(conv_b) != 18 or len(prelu) != 17:
            raise ValueError("SRVGGNetCompact(num_conv=16) has 18 convolutions and 17 PReLUs")
        shapes = [(64, 3, 3, 3)] + [(64, 64, 3, 3)] * 16 + [(3 * scale * scale, 64, 3, 3)]
        keep = []
        for k in range(18):
            w = np.ascontiguousarray(conv_w[k], np.float32)
            b = np.ascontiguousarray(conv_b[k], np.float32)
            if w.shape != shapes[k] or b.shape != (shapes[k][0],):
                raise ValueError(f"convolution {k}: expected weight {shapes[k]}, got {w.shape} / bias {b.shape}")
            keep += [w, b]
        slopes = [np.ascontiguousarray(a, np.float32) for a in prelu]
        if any(a.shape != (64,) for a in slopes):
            raise ValueError("PReLU slopes must have shape (64,)")
        arr = C.c_void_p * 18
        pw = arr(*[keep[2 * k].ctypes.data for k in range(18)])
        pb = arr(*[keep[2 * k + 1].ctypes.data for k in range(18)])
        ps = arr(*([a.ctypes.data for a in slopes] + [None]))
        h = C.c_void_p()
        _check(_lib.load().reve_model_from_arrays(scale, pw, pb, ps, C.byref(h)))
        return cls(h)

    @classmethod
    def from_state_dict(cls, sd, scale: Optional[int] = None) -> "Model":
        """A Real-ESRGAN SRVGGNetCompact state dict (``body.{0,2,..,34}.weight/.bias`` = convolutions,
        ``body.{1,3,..,33}.weight`` = PReLU slopes; values: torch tensors or numpy arrays).  Checkpoints
        wrap it as ``{"params": ...}`` or ``{"params_ema": ...}``; either is unwrapped
        (SURVEY.md section 8(f) row 4)."""
        for key in ("params_ema", "params"):
            if key in sd and hasattr(sd[key], "keys"):
                sd = sd[key]
                break

        def arr(name):
            if name not in sd:
                raise ValueError(f"state dict has no '{name}': not a realesr-animevideov3 (SRVGGNetCompact num_conv=16) checkpoint")
            v = sd[name]
            if hasattr(v, "detach"):
                v = v.detach().cpu().float().numpy()
            return np.asarray(v, np.float32)

        conv_w = [arr(f"body.{2 * k}.weight") for k in range(18)]
        conv_b = [arr(f"body.{2 * k}.bias") for k in range(18)]
        prelu = [arr(f"body.{2 * k + 1}.weight").reshape(-1) for k in range(17)]
        out_ch = conv_w[17].shape[0]
        s = {12: 2, 27: 3, 48: 4}.get(out_ch)
        if s is None or (scale is not None and scale != s):
            raise ValueError(f"last convolution has {out_ch} output channels: not a x{scale or '2/3/4'} model")
        return cls.from_arrays(s, conv_w, conv_b, prelu)

    @classmethod
    def from_pth(cls, path: str, scale: Optional[int] = None) -> "Model":
        """``realesr-animevideov3.pth`` and friends (needs torch to unpickle the checkpoint)."""
        import torch
        return cls.from_state_dict(torch.load(path, map_location="cpu", weights_only=True), scale)

    @classmethod
    def for_scale(cls, scale: int, model_dir: str = "models", allow_random: bool = False, seed: int = 0) -> "Model":
        """The model the reference's CLI intends for ``-s scale``:
        ``<model_dir>/realesr-animevideov3-x{scale}.param|.bin`` (or ``.pth``).  A missing model is an error, as it
        is for the spawned upscaler -- a segment of noise frames with "done" lines and exit code 0 would be worse
        than no output.  ``allow_random=True`` is the explicit opt-in (tests, benches: the weight files are not
        available offline) for the seeded random init of the same architecture."""
        base = os.path.join(model_dir, f"realesr-animevideov3-x{scale}")
        if os.path.exists(base + ".param") and os.path.exists(base + ".bin"):
            return cls.load_ncnn(base + ".param", base + ".bin")
        if os.path.exists(base + ".pth"):
            return cls.from_pth(base + ".pth", scale)
        if not allow_random:
            raise ReveError(-4, f"model files {base}.param/.bin not found (model_dir is resolved against the current "
                                f"directory, {os.getcwd()}); pass allow_random=True for random-init weights")
        return cls.random(scale, seed)

    def save_ncnn(self, param_path: str, bin_path: str, fp16: bool = True) -> None:
        _check(_lib.load().reve_model_save_ncnn(self._h, param_path.encode(), bin_path.encode(), int(fp16)))

    @property
    def scale(self) -> int:
        s = C.c_int()
        _check(_lib.load().reve_model_info(self._h, C.byref(s), None, None))
        return s.value

    def close(self):
        if self._h:
            _lib.load().reve_model_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def geometry(in_w: int, in_h: int, scale: int, tile: int, prepad: int):
    """Canvas tables (host only): (canvas_w, canvas_h, src_x, out_x, src_y, out_y)."""
    lib = _lib.load()
    cw, ch = C.c_int(), C.c_int()
    _check(lib.reve_geometry(in_w, in_h, scale, tile, prepad, C.byref(cw), C.byref(ch), None, None, None, None, 0))
    n = max(cw.value, ch.value)
    arrs = [np.zeros(n, np.int32) for _ in range(4)]
    _check(lib.reve_geometry(in_w, in_h, scale, tile, prepad, C.byref(cw), C.byref(ch),
                             *[a.ctypes.data for a in arrs], n))
    return (cw.value, ch.value, arrs[0][:cw.value].copy(), arrs[1][:cw.value].copy(),
            arrs[2][:ch.value].copy(), arrs[3][:ch.value].copy())


def launch_plan(in_w: int, in_h: int, scale: int, tile: int = 200, prepad: int = 10) -> dict:
    """Launch structure chosen for a frame size (host only): body layers per launch, strip width, strips, launches."""
    lib = _lib.load()
    v = [C.c_int() for _ in range(4)]
    _check(lib.reve_launch_plan(in_w, in_h, scale, tile, prepad, *[C.byref(x) for x in v]))
    return {"layers_per_launch": v[0].value, "strip_px": v[1].value, "n_strips": v[2].value,
            "launches_per_batch": v[3].value}


class Upscaler:
    """One GPU context for a fixed input frame size.  tile=200, prepad=10 are the values the
    reference's spawned upscaler uses; tile=0 is the whole-frame variant."""

    def __init__(self, model: Model, in_w: int, in_h: int, tile: int = 200, prepad: int = 10,
                 device: int = 0, ring_depth: int = 3, *, shared_device: bool = False, layers_per_launch: int = 0,
                 max_batch: int = 0, debug_flags: int = 0, debug_grid: int = 0, trace: int = 0, trace_launch: int = 1,
                 trace_chain: int = 0):
        """Keyword options map one to one onto ``reve_ctx_options`` (include/reve_cuda.h); the library reads no
        environment variables."""
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.model = model
        opt = _lib.reve_ctx_options()
        opt.struct_size = C.sizeof(_lib.reve_ctx_options)
        opt.flags = CTX_SHARED_DEVICE if shared_device else 0
        opt.layers_per_launch, opt.max_batch = layers_per_launch, max_batch
        opt.debug_flags, opt.debug_grid = debug_flags, debug_grid
        opt.trace, opt.trace_launch, opt.trace_chain = trace, trace_launch, trace_chain
        _check(self._lib.reve_ctx_create_ex(device, model._h, in_w, in_h, tile, prepad, ring_depth, C.byref(opt),
                                            C.byref(self._h)))
        v = [C.c_int() for _ in range(5)]
        _check(self._lib.reve_ctx_info(self._h, *[C.byref(x) for x in v]), self._h)
        self.in_w, self.in_h, self.out_w, self.out_h, self.scale = (x.value for x in v)
        self.ring_depth = ring_depth
        self.device = device
        self._pinned: List[int] = []

    def launch_info(self) -> dict:
        """Launch structure in use: body layers per launch, frames per launch set, CTAs, cooperative launch."""
        v = [C.c_int() for _ in range(4)]
        _check(self._lib.reve_ctx_launch_info(self._h, *[C.byref(x) for x in v]), self._h)
        return {"layers_per_launch": v[0].value, "batch": v[1].value, "grid": v[2].value, "cooperative": bool(v[3].value)}

    # -- pinned host buffers ----------------------------------------------------------------
    def pinned(self, shape: Sequence[int]) -> np.ndarray:
        n = int(np.prod(shape))
        p = C.c_void_p()
        _check(self._lib.reve_host_alloc(n, C.byref(p)))
        self._pinned.append(p.value)
        buf = (C.c_uint8 * n).from_address(p.value)
        return np.frombuffer(buf, dtype=np.uint8).reshape(shape)

    # -- frame API ----------------------------------------------------------------------------
    def submit(self, frame: np.ndarray, out: np.ndarray, tag: int = 0) -> None:
        if frame.dtype != np.uint8 or frame.shape != (self.in_h, self.in_w, 3):
            raise ValueError(f"frame must be u8 [{self.in_h},{self.in_w},3]")
        if out.dtype != np.uint8 or out.shape != (self.out_h, self.out_w, 3):
            raise ValueError(f"out must be u8 [{self.out_h},{self.out_w},3]")
        if frame.strides[1:] != (3, 1) or out.strides[1:] != (3, 1):
            raise ValueError("pixels must be packed RGB")
        _check(self._lib.reve_submit(self._h, frame.ctypes.data, frame.strides[0], out.ctypes.data,
                                     out.strides[0], tag), self._h)

    def wait(self) -> int:
        tag = C.c_uint64()
        _check(self._lib.reve_wait(self._h, C.byref(tag)), self._h)
        return tag.value

    def sync(self) -> None:
        _check(self._lib.reve_sync(self._h), self._h)

    def upscale(self, frame: np.ndarray) -> np.ndarray:
        """Synchronous single frame (tests)."""
        frame = np.ascontiguousarray(frame)
        out = np.empty((self.out_h, self.out_w, 3), np.uint8)
        self.submit(frame, out, 0)
        self.wait()
        return out

    # -- yuv420p10le output (what the reference's x265 encode stage consumes) -----------------------
    def set_output_format(self, fmt: int) -> None:
        """FMT_RGB24 (default), FMT_YUV420P10LE_BT601 (swscale's matrix for untagged RGB, i.e. what the
        reference's `-pix_fmt yuv420p10le` produces) or FMT_YUV420P10LE_BT709."""
        _check(self._lib.reve_ctx_set_output_format(self._h, fmt), self._h)
        self.out_format = fmt

    def output_layout(self) -> Tuple[int, int]:
        """(minimal row stride in bytes, bytes per frame at that stride) for the current format."""
        a, b = C.c_size_t(), C.c_size_t()
        _check(self._lib.reve_ctx_output_layout(self._h, C.byref(a), C.byref(b)), self._h)
        return a.value, b.value

    def submit_raw(self, frame: np.ndarray, out: np.ndarray, out_stride: int, tag: int = 0) -> None:
        """`out`: flat u8 buffer laid out as include/reve_cuda.h describes for the current format."""
        if frame.dtype != np.uint8 or frame.shape != (self.in_h, self.in_w, 3) or frame.strides[1:] != (3, 1):
            raise ValueError(f"frame must be packed u8 [{self.in_h},{self.in_w},3]")
        _check(self._lib.reve_submit(self._h, frame.ctypes.data, frame.strides[0], out.ctypes.data, out_stride, tag), self._h)

    def upscale_yuv(self, frame: np.ndarray, out_stride: int = 0):
        """Synchronous single frame in the current YUV format: (Y [H,W], U [ceil(H/2),ceil(W/2)], V) u16."""
        stride, _ = self.output_layout()
        stride = max(stride, out_stride)
        ch, cw = (self.out_h + 1) // 2, (self.out_w + 1) // 2
        buf = np.full(stride * self.out_h + 2 * (stride // 2) * ch, 0xAB, np.uint8)
        self.submit_raw(np.ascontiguousarray(frame), buf, stride, 0)
        self.wait()
        y = buf[:stride * self.out_h].reshape(self.out_h, stride)[:, :2 * self.out_w].copy().view("<u2")
        u = buf[stride * self.out_h:][:(stride // 2) * ch].reshape(ch, stride // 2)
        v = buf[stride * self.out_h + (stride // 2) * ch:].reshape(ch, stride // 2)
        self._last_raw = buf
        return y, u[:, :2 * cw].copy().view("<u2"), v[:, :2 * cw].copy().view("<u2")

    def upscale_many(self, frames: Iterable[np.ndarray], outs: Iterable[np.ndarray],
                     on_done: Optional[Callable[[int], None]] = None) -> int:
        """Pipelined: keeps up to ring_depth frames in flight (H2D / compute / D2H overlapped)."""
        inflight = 0
        n = 0
        for i, (f, o) in enumerate(zip(frames, outs)):
            if inflight == self.ring_depth:
                t = self.wait()
                inflight -= 1
                if on_done:
                    on_done(t)
            self.submit(f, o, i)
            inflight += 1
            n += 1
        while inflight:
            t = self.wait()
            inflight -= 1
            if on_done:
                on_done(t)
        return n

    def upscale_device(self, d_in_ptr: int, d_out_ptr: int, n_frames: int) -> None:
        _check(self._lib.reve_upscale_device(self._h, d_in_ptr, d_out_ptr, n_frames), self._h)

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        _check(self._lib.reve_ctx_stream(self._h, C.byref(s)), self._h)
        return s.value or 0

    def set_profiling(self, on: bool) -> None:
        _check(self._lib.reve_ctx_set_profiling(self._h, int(on)), self._h)

    def profile(self, reset: bool = True) -> dict:
        p = _lib.reve_profile()
        _check(self._lib.reve_ctx_get_profile(self._h, C.byref(p), int(reset)), self._h)
        return {k: getattr(p, k) for k, _ in p._fields_}

    def debug_features(self, frame: np.ndarray, layer: int) -> np.ndarray:
        """fp16 feature canvas after `layer` conv+PReLU stages, as float32 [CH, CW, 64]."""
        frame = np.ascontiguousarray(frame)
        cw, ch = C.c_int(), C.c_int()
        self._lib.reve_debug_features(self._h, None, 0, 0, None, 0, C.byref(cw), C.byref(ch))
        out = np.empty((ch.value, cw.value, 64), np.float32)
        _check(self._lib.reve_debug_features(self._h, frame.ctypes.data, frame.strides[0], layer,
                                             out.ctypes.data, out.size, C.byref(cw), C.byref(ch)), self._h)
        return out

    # -- stand-alone conversion kernels (pack.cu; test / measurement hooks) ------------------------------------
    def debug_unpack(self, frame: np.ndarray, reps: int = 0):
        """u8 frame -> (x/255 as fp16 on the canvas, returned float32 [CH, CW, 3]; kernel ms averaged over reps)."""
        frame = np.ascontiguousarray(frame)
        cw, ch = C.c_int(), C.c_int()
        self._lib.reve_debug_features(self._h, None, 0, 0, None, 0, C.byref(cw), C.byref(ch))
        out = np.empty((ch.value, cw.value, 3), np.float32)
        ms = C.c_float()
        _check(self._lib.reve_debug_unpack(self._h, frame.ctypes.data, frame.strides[0], out.ctypes.data, out.size, reps,
                                           C.byref(ms)), self._h)
        return out, ms.value

    def debug_pack(self, y: np.ndarray, reps: int = 0):
        """float32 [CH*s, CW*s, 3] network output at canvas geometry -> (cropped u8 frame [H*s, W*s, 3], kernel ms)."""
        y = np.ascontiguousarray(y, np.float32)
        out = np.empty((self.out_h, self.out_w, 3), np.uint8)
        ms = C.c_float()
        _check(self._lib.reve_debug_pack(self._h, y.ctypes.data, y.size, out.ctypes.data, out.strides[0], reps, C.byref(ms)), self._h)
        return out, ms.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.reve_ctx_destroy(self._h)
            self._h = None
        for p in getattr(self, "_pinned", []):
            self._lib.reve_host_free(p)
        self._pinned = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
# Segment-level mirror of Video::upscale_segment
# ----------------------------------------------------------------------------------------------
def _read_frame(path: str) -> np.ndarray:
    if path.endswith(".npy"):
        return np.load(path)
    import cv2  # host-side decode only (the reference decodes PNG on the host as well)
    img = cv2.imread(path, cv2.IMREAD_COLOR)
    if img is None:
        raise IOError(f"cannot decode {path}")
    return np.ascontiguousarray(img[:, :, ::-1])


def _write_frame(path: str, rgb: np.ndarray) -> None:
    if path.endswith(".npy"):
        np.save(path, rgb)
        return
    import cv2
    if not cv2.imwrite(path, np.ascontiguousarray(rgb[:, :, ::-1])):
        raise IOError(f"cannot encode {path}")


def upscale_segment(input_dir: str, output_dir: str, scale: int, model: Optional[Model] = None,
                    tile: int = 200, prepad: int = 10, device: int = 0, fmt: str = "png",
                    progress=None, model_dir: str = "models", allow_random: bool = False) -> int:
    """Drop-in for the process spawned by Video::upscale_segment (reference
    reve-shared/src/lib.rs:134-147: ``-i input_dir -o output_dir -n realesr-animevideov3-x2 -s
    scale -f png -v``): every frame file of ``input_dir`` (sorted by name) is upscaled into
    ``output_dir`` under the same stem with extension ``fmt``; with ``progress`` (a text stream)
    one ``"<in> -> <out> done"`` line is written per finished frame, which is what
    reve-cli/src/main.rs:265-273 counts.  Unlike the reference, failures raise instead of being
    ignored.  Returns the number of frames written."""
    names = sorted(n for n in os.listdir(input_dir)
                   if n.lower().endswith((".png", ".jpg", ".jpeg", ".webp", ".npy")))
    if not names:
        return 0
    os.makedirs(output_dir, exist_ok=True)
    own_model = model is None
    if own_model:
        model = Model.for_scale(scale, model_dir, allow_random=allow_random)
    if model.scale != scale:
        raise ValueError(f"model is x{model.scale} but scale {scale} was requested")
    first = _read_frame(os.path.join(input_dir, names[0]))
    h, w = first.shape[:2]
    # frames smaller than the pre-pad: upstream's reflect-101 reads out of bounds there (undefined); we pad as far as
    # the rule is defined
    prepad = min(prepad, min(w, h) - 1)
    up = Upscaler(model, w, h, tile=tile, prepad=prepad, device=device, ring_depth=3)
    try:
        ins = [up.pinned((h, w, 3)) for _ in range(up.ring_depth)]
        outs = [up.pinned((up.out_h, up.out_w, 3)) for _ in range(up.ring_depth)]
        pending: List[Tuple[int, str, str]] = []
        done = 0

        def retire():
            nonlocal done
            up.wait()
            slot, src, dst = pending.pop(0)
            _write_frame(dst, outs[slot])
            done += 1
            if progress is not None:
                progress.write(f"{src} -> {dst} done\n")
                progress.flush()

        for i, name in enumerate(names):
            slot = i % up.ring_depth
            if len(pending) == up.ring_depth:
                retire()
            frame = first if i == 0 else _read_frame(os.path.join(input_dir, name))
            if frame.shape != (h, w, 3):
                raise ValueError(f"{name}: frame size differs from the first frame of the segment")
            ins[slot][...] = frame
            src = os.path.join(input_dir, name)
            dst = os.path.join(output_dir, os.path.splitext(name)[0] + "." + fmt)
            up.submit(ins[slot], outs[slot], i)
            pending.append((slot, src, dst))
        while pending:
            retire()
        return done
    finally:
        up.close()
        if own_model:
            model.close()
