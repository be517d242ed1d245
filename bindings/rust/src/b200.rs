//! `liblrge/src/b200.rs` -- the B200 path behind liblrge's `Estimate` trait (cargo feature `b200`).
//!
//! Replaces, for `TwoSetStrategy` and `AvaStrategy`, everything between "temp FASTA written" and
//! "`Vec<f32>` estimates returned": `AlignerWrapper::new` + `align_reads` / `align_reads_inverse`
//! (liblrge/src/twoset.rs:596-604, liblrge/src/ava.rs:378-380) become ONE call into liblrge_b200.so
//! (include/lrge_b200.h).  `Estimate`, the Builders, `split_fastq` / `subsample_reads` and the CLI stay untouched;
//! `minimap2-sys`, `AlignerWrapper` and `thread_buf` are not used on this path.
use std::ffi::CStr;
use std::os::raw::{c_char, c_int};
use std::path::Path;

use needletail::parse_fastx_file;

use crate::{LrgeError, Platform};

/// `lrge_readset`: a temp FASTA as flat buffers; read order = rid order.
#[repr(C)]
pub struct LrgeReadset {
    bases: *const u8,
    off: *const u64,
    names: *const c_char,
    name_off: *const u32,
    n: u32,
}

/// `lrge_opts`: the Builder fields that reach the hot path (twoset/builder.rs:22-39, ava/builder.rs:18-32).
#[repr(C)]
#[derive(Default)]
pub struct LrgeOpts {
    platform: i32,
    remove_internal: i32,
    max_overhang_ratio: f32,
    use_min_ref: i32,
    shard_index: i32,
    shard_count: i32,
    reserved: [i32; 6],
}

/// `lrge_stats` is only ever passed as NULL from here.
#[repr(C)]
pub struct LrgeStats {
    _private: [u8; 0],
}

pub enum LrgeCtx {}

extern "C" {
    fn lrge_opts_default(o: *mut LrgeOpts);
    fn lrge_gpu_create(device: c_int, out: *mut *mut LrgeCtx) -> c_int;
    fn lrge_gpu_destroy(ctx: *mut LrgeCtx);
    fn lrge_gpu_last_error(ctx: *const LrgeCtx) -> *const c_char;
    fn lrge_twoset_generate_estimates(
        ctx: *mut LrgeCtx,
        t: *const LrgeReadset,
        q: *const LrgeReadset,
        o: *const LrgeOpts,
        estimates: *mut f32,
        no_mapping_count: *mut u32,
        stats: *mut LrgeStats,
    ) -> c_int;
    fn lrge_ava_generate_estimates(
        ctx: *mut LrgeCtx,
        r: *const LrgeReadset,
        o: *const LrgeOpts,
        estimates: *mut f32,
        no_mapping_count: *mut u32,
        stats: *mut LrgeStats,
    ) -> c_int;
    // one process per GPU: rank 0 draws the id, every rank joins, then each rank passes ITS block of the targets
    fn lrge_gpu_comm_unique_id(out: *mut u8, size: usize) -> c_int;
    fn lrge_gpu_comm_init(ctx: *mut LrgeCtx, rank: c_int, world: c_int, unique_id: *const u8, size: usize) -> c_int;
    fn lrge_gpu_twoset_counts_shard(
        ctx: *mut LrgeCtx,
        t_block: *const LrgeReadset,
        q: *const LrgeReadset,
        o: *const LrgeOpts,
        counts: *mut u32,
        had_mapping: *mut u8,
        stats: *mut LrgeStats,
    ) -> c_int;
}

/// The temp FASTA as flat buffers (what `lrge_readset` points into).  The read id is the header up to the first
/// whitespace, exactly what liblrge writes into the temp FASTA (io.rs:192-205, twoset.rs:157-186).
struct Flat {
    bases: Vec<u8>,
    off: Vec<u64>,
    names: Vec<u8>,
    name_off: Vec<u32>,
}

impl Flat {
    fn load(path: &Path) -> crate::Result<Self> {
        let mut f = Flat { bases: vec![], off: vec![0], names: vec![], name_off: vec![0] };
        let mut rdr = parse_fastx_file(path).map_err(|e| LrgeError::FastqParseError(e.to_string()))?;
        while let Some(rec) = rdr.next() {
            let rec = rec.map_err(|e| LrgeError::FastqParseError(e.to_string()))?;
            f.bases.extend_from_slice(&rec.seq());
            f.off.push(f.bases.len() as u64);
            let id = rec.id();
            let end = id.iter().position(|b| b.is_ascii_whitespace()).unwrap_or(id.len());
            f.names.extend_from_slice(&id[..end]);
            f.name_off.push(f.names.len() as u32);
        }
        Ok(f)
    }

    fn n(&self) -> usize {
        self.off.len() - 1
    }

    fn view(&self) -> LrgeReadset {
        LrgeReadset {
            bases: self.bases.as_ptr(),
            off: self.off.as_ptr(),
            names: self.names.as_ptr() as *const c_char,
            name_off: self.name_off.as_ptr(),
            n: self.n() as u32,
        }
    }
}

/// Owns an `lrge_ctx`; one per process per GPU.
struct Ctx(*mut LrgeCtx);

impl Ctx {
    fn new(device: c_int) -> crate::Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { lrge_gpu_create(device, &mut ctx) };
        if rc != 0 {
            return Err(map_err(rc, std::ptr::null()));
        }
        Ok(Ctx(ctx))
    }
}

impl Drop for Ctx {
    fn drop(&mut self) {
        unsafe { lrge_gpu_destroy(self.0) }
    }
}

/// lrge_status -> LrgeError (include/lrge_b200.h mirrors liblrge/src/error.rs:6-33).
fn map_err(code: c_int, ctx: *const LrgeCtx) -> LrgeError {
    let msg = unsafe { CStr::from_ptr(lrge_gpu_last_error(ctx)) }.to_string_lossy().into_owned();
    match code {
        9 => LrgeError::DuplicateReadIdentifier(msg),
        2 => LrgeError::FastqParseError(msg),
        _ => LrgeError::MapError(msg), // 8 = MapError; CUDA / NCCL failures surface as MapError too
    }
}

fn opts(platform: &Platform, remove_internal: bool, max_overhang_ratio: f32, use_min_ref: bool) -> LrgeOpts {
    let mut o = LrgeOpts::default();
    unsafe { lrge_opts_default(&mut o) };
    o.platform = matches!(platform, Platform::PacBio) as i32; // PacBio -> ava-pb, Nanopore -> ava-ont (twoset.rs:591-594)
    o.remove_internal = remove_internal as i32;
    o.max_overhang_ratio = max_overhang_ratio;
    o.use_min_ref = use_min_ref as i32;
    o
}

/// `TwoSetStrategy::generate_estimates` after `split_fastq` (twoset.rs:596-604): per-query estimates and the
/// number of queries without any mapping.  With `use_min_ref` the library picks the direction as twoset.rs:596 does.
pub fn twoset_estimates(
    target: &Path,
    query: &Path,
    platform: &Platform,
    remove_internal: bool,
    max_overhang_ratio: f32,
    use_min_ref: bool,
) -> crate::Result<(Vec<f32>, u32)> {
    let (t, q) = (Flat::load(target)?, Flat::load(query)?);
    let o = opts(platform, remove_internal, max_overhang_ratio, use_min_ref);
    let ctx = Ctx::new(0)?;
    let mut est = vec![0f32; q.n()];
    let mut no_mapping = 0u32;
    let rc = unsafe {
        lrge_twoset_generate_estimates(ctx.0, &t.view(), &q.view(), &o, est.as_mut_ptr(), &mut no_mapping, std::ptr::null_mut())
    };
    if rc == 0 {
        Ok((est, no_mapping))
    } else {
        Err(map_err(rc, ctx.0))
    }
}

/// `AvaStrategy::generate_estimates` after `subsample_reads` (ava.rs:378-380).
pub fn ava_estimates(
    reads: &Path,
    platform: &Platform,
    remove_internal: bool,
    max_overhang_ratio: f32,
) -> crate::Result<(Vec<f32>, u32)> {
    let r = Flat::load(reads)?;
    let o = opts(platform, remove_internal, max_overhang_ratio, false);
    let ctx = Ctx::new(0)?;
    let mut est = vec![0f32; r.n()];
    let mut no_mapping = 0u32;
    let rc = unsafe { lrge_ava_generate_estimates(ctx.0, &r.view(), &o, est.as_mut_ptr(), &mut no_mapping, std::ptr::null_mut()) };
    if rc == 0 {
        Ok((est, no_mapping))
    } else {
        Err(map_err(rc, ctx.0))
    }
}

/// Communicator id for a multi-GPU run (128 bytes): rank 0 calls this and hands the bytes to the other ranks by any
/// means (MPI, a file, a socket).
pub fn comm_unique_id() -> crate::Result<[u8; 128]> {
    let mut id = [0u8; 128];
    let rc = unsafe { lrge_gpu_comm_unique_id(id.as_mut_ptr(), id.len()) };
    if rc == 0 {
        Ok(id)
    } else {
        Err(map_err(rc, std::ptr::null()))
    }
}

/// Target sets that do not fit one host (README.md:173-175): every rank (one process per GPU) passes only ITS block of the
/// target FASTA and the whole query FASTA, and gets the complete per-query overlap counts and `had_mapping` flags back --
/// the inputs of `per_read_estimate` (estimate.rs:129-157; `avg_target_len` and the target count are taken over all blocks
/// by the caller).  Target read ids must be unique across blocks.
pub fn twoset_counts_shard(
    device: c_int,
    rank: c_int,
    world: c_int,
    unique_id: &[u8; 128],
    target_block: &Path,
    query: &Path,
    platform: &Platform,
    remove_internal: bool,
    max_overhang_ratio: f32,
) -> crate::Result<(Vec<u32>, Vec<u8>)> {
    let (t, q) = (Flat::load(target_block)?, Flat::load(query)?);
    let mut o = opts(platform, remove_internal, max_overhang_ratio, false);
    o.shard_index = rank;
    o.shard_count = world;
    let ctx = Ctx::new(device)?;
    let rc = unsafe { lrge_gpu_comm_init(ctx.0, rank, world, unique_id.as_ptr(), unique_id.len()) };
    if rc != 0 {
        return Err(map_err(rc, ctx.0));
    }
    let mut counts = vec![0u32; q.n()];
    let mut had = vec![0u8; q.n()];
    let rc = unsafe {
        lrge_gpu_twoset_counts_shard(ctx.0, &t.view(), &q.view(), &o, counts.as_mut_ptr(), had.as_mut_ptr(), std::ptr::null_mut())
    };
    if rc == 0 {
        Ok((counts, had))
    } else {
        Err(map_err(rc, ctx.0))
    }
}
