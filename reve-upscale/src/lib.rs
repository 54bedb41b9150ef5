//! In-process replacement for the `realesrgan-ncnn-vulkan` child that
//! `Video::upscale_segment` spawns (reve-shared/src/lib.rs:129-155).
//!
//! `upscale_segment(..)` keeps the shape the caller relies on (reve-cli/src/main.rs:262-273): it
//! returns a `BufRead` that yields one line containing `done` per finished frame and reaches EOF
//! when the segment is complete.  The work runs on a worker thread that drives one `reve_ctx`
//! (one GPU) through the C ABI of include/reve_cuda.h.  Unlike the reference, a failure is
//! surfaced: the last line is `error: ...` and `SegmentHandle::join` returns `Err`.
//!
//! NOTE: written against the C ABI but not compiled here (no Rust toolchain in the build image).
use std::ffi::{c_char, c_int, c_void, CStr, CString};
use std::io::{BufRead, BufReader, Write};
use std::os::unix::net::UnixStream;
use std::path::{Path, PathBuf};
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
    /// `models/realesr-animevideov3-x{scale}.param|.bin` relative to the exe dir (main.rs:109),
    /// or the seeded random init when the files are absent.
    pub fn for_scale(model_dir: &Path, scale: u8) -> Result<Model, String> {
        let stem = model_dir.join(format!("realesr-animevideov3-x{}", scale));
        let (p, b) = (stem.with_extension("param"), stem.with_extension("bin"));
        let mut m = std::ptr::null_mut();
        let rc = if p.exists() && b.exists() {
            let (p, b) = (CString::new(p.to_str().unwrap()).unwrap(), CString::new(b.to_str().unwrap()).unwrap());
            unsafe { reve_model_load_ncnn(p.as_ptr(), b.as_ptr(), &mut m) }
        } else {
            unsafe { reve_model_random(scale as c_int, 1234, &mut m) }
        };
        if rc != 0 { Err(last_error(std::ptr::null())) } else { Ok(Model(m, scale)) }
    }
}
impl Drop for Model { fn drop(&mut self) { unsafe { reve_model_free(self.0) } } }

pub struct SegmentHandle { worker: JoinHandle<Result<usize, String>> }
impl SegmentHandle {
    pub fn join(self) -> Result<usize, String> { self.worker.join().map_err(|_| "upscale worker panicked".to_string())? }
}

/// Drop-in for `Video::upscale_segment`: frames `input_dir/frame%08d.png` -> `output_dir/` at `scale`.
pub fn upscale_segment(model: std::sync::Arc<Model>, device: i32, input_dir: PathBuf, output_dir: PathBuf)
    -> std::io::Result<(BufReader<UnixStream>, SegmentHandle)> {
    std::fs::create_dir(&output_dir)?;                       // lib.rs:130-132
    let (rx, mut tx) = UnixStream::pair()?;                  // stands in for the child's stderr pipe
    let worker = std::thread::spawn(move || -> Result<usize, String> {
        let res = run_segment(&model, device, &input_dir, &output_dir, &mut tx);
        if let Err(e) = &res { let _ = writeln!(tx, "error: {}", e); }
        res
    });
    Ok((BufReader::new(rx), SegmentHandle { worker }))
}

fn run_segment(model: &Model, device: i32, input_dir: &Path, output_dir: &Path, progress: &mut UnixStream) -> Result<usize, String> {
    let mut names: Vec<PathBuf> = std::fs::read_dir(input_dir).map_err(|e| e.to_string())?
        .filter_map(|e| e.ok().map(|e| e.path())).filter(|p| p.extension().map_or(false, |x| x == "png")).collect();
    names.sort();
    if names.is_empty() { return Ok(0); }
    let (w, h, first) = read_png(&names[0])?;
    let s = model.1 as usize;
    let mut ctx = std::ptr::null_mut();
    if unsafe { reve_ctx_create(device, model.0, w as c_int, h as c_int, 200, 10, 3, &mut ctx) } != 0 {
        return Err(last_error(std::ptr::null()));
    }
    let (in_bytes, out_bytes) = (w * h * 3, w * h * 3 * s * s);
    let mut bufs: Vec<(*mut u8, *mut u8)> = Vec::new();
    for _ in 0..3 {
        let (mut a, mut b) = (std::ptr::null_mut(), std::ptr::null_mut());
        unsafe { reve_host_alloc(in_bytes, &mut a); reve_host_alloc(out_bytes, &mut b); }
        bufs.push((a as *mut u8, b as *mut u8));
    }
    let mut pending: std::collections::VecDeque<(usize, PathBuf, PathBuf)> = Default::default();
    let mut done = 0usize;
    let mut retire = |pending: &mut std::collections::VecDeque<(usize, PathBuf, PathBuf)>, done: &mut usize| -> Result<(), String> {
        let mut tag = 0u64;
        if unsafe { reve_wait(ctx, &mut tag) } != 0 { return Err(last_error(ctx)); }
        let (slot, src, dst) = pending.pop_front().unwrap();
        let out = unsafe { std::slice::from_raw_parts(bufs[slot].1, out_bytes) };
        write_png(&dst, out, w * s, h * s)?;
        *done += 1;
        let _ = writeln!(progress, "{} -> {} done", src.display(), dst.display());   // counted at main.rs:269
        Ok(())
    };
    let mut result = Ok(());
    for (i, src) in names.iter().enumerate() {
        let slot = i % 3;
        if pending.len() == 3 { if let Err(e) = retire(&mut pending, &mut done) { result = Err(e); break; } }
        let data = if i == 0 { first.clone() } else { let (fw, fh, d) = read_png(src)?; if (fw, fh) != (w, h) { result = Err("frame size changed".into()); break; } d };
        unsafe { std::ptr::copy_nonoverlapping(data.as_ptr(), bufs[slot].0, in_bytes); }
        let dst = output_dir.join(src.file_name().unwrap());
        if unsafe { reve_submit(ctx, bufs[slot].0, w * 3, bufs[slot].1, w * s * 3, i as u64) } != 0 { result = Err(last_error(ctx)); break; }
        pending.push_back((slot, src.clone(), dst));
    }
    while result.is_ok() && !pending.is_empty() { result = retire(&mut pending, &mut done); }
    for (a, b) in bufs { unsafe { reve_host_free(a as *mut c_void); reve_host_free(b as *mut c_void); } }
    unsafe { reve_ctx_destroy(ctx) };
    result.map(|_| done)
}

fn read_png(p: &Path) -> Result<(usize, usize, Vec<u8>), String> {
    let dec = png::Decoder::new(std::fs::File::open(p).map_err(|e| e.to_string())?);
    let mut r = dec.read_info().map_err(|e| e.to_string())?;
    let mut buf = vec![0; r.output_buffer_size()];
    let info = r.next_frame(&mut buf).map_err(|e| e.to_string())?;
    if info.color_type != png::ColorType::Rgb || info.bit_depth != png::BitDepth::Eight { return Err(format!("{}: expected 8-bit RGB", p.display())); }
    buf.truncate(info.buffer_size());
    Ok((info.width as usize, info.height as usize, buf))
}

fn write_png(p: &Path, rgb: &[u8], w: usize, h: usize) -> Result<(), String> {
    let mut enc = png::Encoder::new(std::io::BufWriter::new(std::fs::File::create(p).map_err(|e| e.to_string())?), w as u32, h as u32);
    enc.set_color(png::ColorType::Rgb);
    enc.set_depth(png::BitDepth::Eight);
    enc.set_compression(png::Compression::Fast);
    enc.write_header().and_then(|mut wr| wr.write_image_data(rgb)).map_err(|e| e.to_string())
}
