//! In-process replacement for the `realesrgan-ncnn-vulkan` child that
//! `Video::upscale_segment` spawns (reve-shared/src/lib.rs:129-155).
//!
//! `upscale_segment(..)` keeps the shape the caller relies on (reve-cli/src/main.rs:262-273): it
//! returns a `BufRead` that yields one line containing `done` per finished frame and reaches EOF
//! when the segment is complete.  The work runs on a worker thread that drives one `reve_ctx`
//! (one GPU) through the C ABI of include/reve_cuda.h.  Unlike the reference, a failure is
//! surfaced: the last line is `error: ...` and `SegmentHandle::join` returns `Err`.
//!
//! Resource discipline: the context and the pinned buffers are owned by RAII guards (`Ctx`,
//! `Pinned`), declared so that the context is dropped FIRST -- `reve_ctx_destroy` synchronises the
//! context's streams, so no DMA can still target a pinned buffer when it is freed, whichever `?`
//! ends `run_segment` early.
//!
//! NOTE: written against the C ABI but not compiled here (no Rust toolchain in the build image);
//! the same behaviour is compiled and tested in reve_b200/host/reve_upscale.cpp and reve_b200/upscaler.py.
use std::ffi::{c_char, c_int, c_void, CStr, CString};
use std::io::{BufRead, BufReader, Read};
use std::path::{Path, PathBuf};
use std::sync::mpsc::{channel, Receiver, Sender};
use std::thread::JoinHandle;

#[repr(C)] pub struct ReveModel { _p: [u8; 0] }
#[repr(C)] pub struct ReveCtx { _p: [u8; 0] }

extern "C" {
    fn reve_last_error(ctx: *const ReveCtx) -> *const c_char;
    fn reve_model_load_ncnn(param: *const c_char, bin: *const c_char, out: *mut *mut ReveModel) -> c_int;
    fn reve_model_random(scale: c_int, seed: u64, out: *mut *mut ReveModel) -> c_int;
    fn reve_model_free(m: *mut ReveModel);
    fn reve_ctx_create(device: c_int, m: *const ReveModel, in_w: c_int, in_h: c_int, tile: c_int, prepad: c_int,
                       ring_depth: c_int, out: *mut *mut ReveCtx) -> c_int;
    fn reve_ctx_destroy(ctx: *mut ReveCtx);
    fn reve_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    fn reve_host_free(p: *mut c_void);
    fn reve_submit(ctx: *mut ReveCtx, rgb_in: *const u8, in_stride: usize, rgb_out: *mut u8, out_stride: usize, tag: u64) -> c_int;
    fn reve_wait(ctx: *mut ReveCtx, tag: *mut u64) -> c_int;
    fn reve_sync(ctx: *mut ReveCtx) -> c_int;
    #[allow(dead_code)]
    fn reve_ctx_set_output_format(ctx: *mut ReveCtx, format: c_int) -> c_int;
    #[allow(dead_code)]
    fn reve_ctx_output_layout(ctx: *const ReveCtx, min_stride: *mut usize, frame_bytes: *mut usize) -> c_int;
    #[allow(dead_code)]
    fn reve_model_from_arrays(scale: c_int, conv_w: *const *const f32, conv_b: *const *const f32,
                              prelu: *const *const f32, out: *mut *mut ReveModel) -> c_int;
}

fn last_error(ctx: *const ReveCtx) -> String {
    unsafe { CStr::from_ptr(reve_last_error(ctx)).to_string_lossy().into_owned() }
}

/// Weights of realesr-animevideov3-x{scale}; `Send + Sync` (immutable after creation).
pub struct Model(*mut ReveModel, pub u8);
unsafe impl Send for Model {}
unsafe impl Sync for Model {}
impl Model {
    /// `models/realesr-animevideov3-x{scale}.param|.bin` relative to the exe dir (main.rs:109).
    /// Missing files are an error, as they are for the spawned upstream binary (a segment of noise
    /// frames with `done` lines would be worse than no output).
    pub fn for_scale(model_dir: &Path, scale: u8) -> Result<Model, String> {
        let stem = model_dir.join(format!("realesr-animevideov3-x{}", scale));
        let (p, b) = (stem.with_extension("param"), stem.with_extension("bin"));
        if !(p.exists() && b.exists()) {
            return Err(format!("model files {}.param/.bin not found", stem.display()));
        }
        let cs = |x: &Path| x.to_str().ok_or_else(|| format!("{}: path is not valid UTF-8", x.display()))
            .and_then(|s| CString::new(s).map_err(|e| e.to_string()));
        let (p, b) = (cs(&p)?, cs(&b)?);
        let mut m = std::ptr::null_mut();
        let rc = unsafe { reve_model_load_ncnn(p.as_ptr(), b.as_ptr(), &mut m) };
        if rc != 0 { Err(last_error(std::ptr::null())) } else { Ok(Model(m, scale)) }
    }
    /// Explicit opt-in (tests, benches): the seeded random init of the same architecture.
    pub fn random(scale: u8, seed: u64) -> Result<Model, String> {
        let mut m = std::ptr::null_mut();
        let rc = unsafe { reve_model_random(scale as c_int, seed, &mut m) };
        if rc != 0 { Err(last_error(std::ptr::null())) } else { Ok(Model(m, scale)) }
    }
}
impl Drop for Model { fn drop(&mut self) { unsafe { reve_model_free(self.0) } } }

/// One `reve_ctx`.  `Send`, not `Sync` (include/reve_cuda.h: a context is not thread-safe).
struct Ctx(*mut ReveCtx);
unsafe impl Send for Ctx {}
impl Ctx {
    fn new(model: &Model, device: i32, w: usize, h: usize, ring: usize) -> Result<Ctx, String> {
        let mut c = std::ptr::null_mut();
        // frames smaller than the pre-pad: upstream's reflect-101 reads out of bounds there
        let prepad = 10.min(w.min(h) as c_int - 1);
        if unsafe { reve_ctx_create(device, model.0, w as c_int, h as c_int, 200, prepad, ring as c_int, &mut c) } != 0 {
            return Err(last_error(std::ptr::null()));
        }
        Ok(Ctx(c))
    }
}
impl Drop for Ctx {
    fn drop(&mut self) {
        // drains every stream of the context before anything it may still DMA into is released
        unsafe { reve_sync(self.0); reve_ctx_destroy(self.0); }
    }
}

/// Pinned host buffer from `reve_host_alloc`.
struct Pinned { p: *mut u8, len: usize }
unsafe impl Send for Pinned {}
impl Pinned {
    fn new(len: usize) -> Result<Pinned, String> {
        let mut p: *mut c_void = std::ptr::null_mut();
        if unsafe { reve_host_alloc(len, &mut p) } != 0 || p.is_null() { return Err(last_error(std::ptr::null())); }
        Ok(Pinned { p: p as *mut u8, len })
    }
    fn as_slice(&self) -> &[u8] { unsafe { std::slice::from_raw_parts(self.p, self.len) } }
    fn as_mut_slice(&mut self) -> &mut [u8] { unsafe { std::slice::from_raw_parts_mut(self.p, self.len) } }
}
impl Drop for Pinned { fn drop(&mut self) { unsafe { reve_host_free(self.p as *mut c_void) } } }

/// The reader side of the progress stream: an mpsc channel behind `Read`, so the caller keeps its
/// `BufRead::lines()` loop (main.rs:265-273) on every platform (the reference is Windows-pathed; no
/// Unix socket pair).  EOF = the worker dropped its sender, i.e. the segment ended.
pub struct ProgressPipe { rx: Receiver<Vec<u8>>, cur: Vec<u8>, pos: usize }
impl Read for ProgressPipe {
    fn read(&mut self, buf: &mut [u8]) -> std::io::Result<usize> {
        while self.pos == self.cur.len() {
            match self.rx.recv() {
                Ok(line) => { self.cur = line; self.pos = 0; }
                Err(_) => return Ok(0),
            }
        }
        let n = buf.len().min(self.cur.len() - self.pos);
        buf[..n].copy_from_slice(&self.cur[self.pos..self.pos + n]);
        self.pos += n;
        Ok(n)
    }
}

pub struct SegmentHandle { worker: JoinHandle<Result<usize, String>> }
impl SegmentHandle {
    pub fn join(self) -> Result<usize, String> { self.worker.join().map_err(|_| "upscale worker panicked".to_string())? }
}

/// Drop-in for `Video::upscale_segment`: frames `input_dir/frame%08d.png` -> `output_dir/` at the model's scale.
pub fn upscale_segment(model: std::sync::Arc<Model>, device: i32, input_dir: PathBuf, output_dir: PathBuf)
    -> std::io::Result<(BufReader<ProgressPipe>, SegmentHandle)> {
    std::fs::create_dir(&output_dir)?;                       // lib.rs:130-132
    let (tx, rx) = channel::<Vec<u8>>();                     // stands in for the child's stderr pipe
    let worker = std::thread::spawn(move || -> Result<usize, String> {
        let res = run_segment(&model, device, &input_dir, &output_dir, &tx);
        if let Err(e) = &res { let _ = tx.send(format!("error: {}\n", e).into_bytes()); }
        res
    });
    Ok((BufReader::new(ProgressPipe { rx, cur: Vec::new(), pos: 0 }), SegmentHandle { worker }))
}

const RING: usize = 3;

fn run_segment(model: &Model, device: i32, input_dir: &Path, output_dir: &Path, progress: &Sender<Vec<u8>>) -> Result<usize, String> {
    let mut names: Vec<PathBuf> = std::fs::read_dir(input_dir).map_err(|e| e.to_string())?
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x.eq_ignore_ascii_case("png"))).collect();
    names.sort();
    if names.is_empty() { return Ok(0); }
    let (w, h, first) = read_png(&names[0])?;
    let s = model.1 as usize;
    let (in_bytes, out_bytes) = (w * h * 3, w * h * 3 * s * s);
    // Declaration order = reverse drop order: `ctx` is declared LAST, so it is dropped (synchronised and destroyed)
    // before the pinned buffers on every exit path, including the `?`s below.
    let mut bufs: Vec<(Pinned, Pinned)> = Vec::with_capacity(RING);
    for _ in 0..RING { bufs.push((Pinned::new(in_bytes)?, Pinned::new(out_bytes)?)); }
    let ctx = Ctx::new(model, device, w, h, RING)?;

    let mut pending: std::collections::VecDeque<(usize, PathBuf, PathBuf)> = Default::default();
    let mut done = 0usize;
    for (i, src) in names.iter().enumerate() {
        let slot = i % RING;
        if pending.len() == RING { retire(&ctx, &bufs, &mut pending, &mut done, (w * s, h * s), progress)?; }
        let data = if i == 0 { first.clone() } else {
            let (fw, fh, d) = read_png(src)?;
            if (fw, fh) != (w, h) { return Err(format!("{}: frame size differs from the first frame of the segment", src.display())); }
            d
        };
        bufs[slot].0.as_mut_slice().copy_from_slice(&data);
        let dst = output_dir.join(src.file_name().unwrap());
        if unsafe { reve_submit(ctx.0, bufs[slot].0.p, w * 3, bufs[slot].1.p, w * s * 3, i as u64) } != 0 { return Err(last_error(ctx.0)); }
        pending.push_back((slot, src.clone(), dst));
    }
    while !pending.is_empty() { retire(&ctx, &bufs, &mut pending, &mut done, (w * s, h * s), progress)?; }
    Ok(done)
}

fn retire(ctx: &Ctx, bufs: &[(Pinned, Pinned)], pending: &mut std::collections::VecDeque<(usize, PathBuf, PathBuf)>, done: &mut usize,
          out_size: (usize, usize), progress: &Sender<Vec<u8>>) -> Result<(), String> {
    let mut tag = 0u64;
    if unsafe { reve_wait(ctx.0, &mut tag) } != 0 { return Err(last_error(ctx.0)); }
    let (slot, src, dst) = pending.pop_front().expect("reve_wait returned a frame that was never submitted");
    write_png(&dst, bufs[slot].1.as_slice(), out_size.0, out_size.1)?;
    *done += 1;
    let _ = progress.send(format!("{} -> {} done\n", src.display(), dst.display()).into_bytes());   // counted at main.rs:269
    Ok(())
}

/// 8-bit PNG of any colour type as packed RGB -- what upstream's loader produces (gray -> RGB, alpha dropped: its
/// separate bicubic alpha path never occurs for ffmpeg-exported frames), and what the C++ and
/// Python drivers accept.
fn read_png(p: &Path) -> Result<(usize, usize, Vec<u8>), String> {
    let mut dec = png::Decoder::new(std::fs::File::open(p).map_err(|e| format!("{}: {}", p.display(), e))?);
    dec.set_transformations(png::Transformations::EXPAND | png::Transformations::STRIP_16);   // palette / <8-bit / 16-bit -> 8-bit
    let mut r = dec.read_info().map_err(|e| format!("{}: {}", p.display(), e))?;
    let mut buf = vec![0; r.output_buffer_size()];
    let info = r.next_frame(&mut buf).map_err(|e| format!("{}: {}", p.display(), e))?;
    buf.truncate(info.buffer_size());
    let (w, h) = (info.width as usize, info.height as usize);
    let rgb = match info.color_type {
        png::ColorType::Rgb => buf,
        png::ColorType::Rgba => buf.chunks_exact(4).flat_map(|q| [q[0], q[1], q[2]]).collect(),
        png::ColorType::Grayscale => buf.iter().flat_map(|&g| [g, g, g]).collect(),
        png::ColorType::GrayscaleAlpha => buf.chunks_exact(2).flat_map(|q| [q[0], q[0], q[0]]).collect(),
        png::ColorType::Indexed => return Err(format!("{}: palette was not expanded", p.display())),
    };
    if rgb.len() != w * h * 3 { return Err(format!("{}: unexpected decoded size", p.display())); }
    Ok((w, h, rgb))
}

fn write_png(p: &Path, rgb: &[u8], w: usize, h: usize) -> Result<(), String> {
    let mut enc = png::Encoder::new(std::io::BufWriter::new(std::fs::File::create(p).map_err(|e| format!("{}: {}", p.display(), e))?), w as u32, h as u32);
    enc.set_color(png::ColorType::Rgb);
    enc.set_depth(png::BitDepth::Eight);
    enc.set_compression(png::Compression::Fast);
    enc.write_header().and_then(|mut wr| wr.write_image_data(rgb)).map_err(|e| format!("{}: {}", p.display(), e))
}
