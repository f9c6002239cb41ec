//! x3::Parameters / Channel / IterChannel / FrameHeader with the reference's names (src/x3.rs).
use crate::error::{check, Result};
use crate::ffi;

pub struct Parameters {
    pub block_len: usize,
    pub blocks_per_frame: usize,
    pub codes: [usize; 3],
    pub thresholds: [usize; 3],
}
impl Parameters {
    pub const MAX_BLOCK_LENGTH: usize = 60;
    pub const DEFAULT_BLOCK_LENGTH: usize = 20;
    pub const DEFAULT_RICE_CODES: [usize; 3] = [0, 1, 3];
    pub const DEFAULT_THRESHOLDS: [usize; 3] = [3, 8, 20];
    pub const DEFAULT_BLOCKS_PER_FRAME: usize = 500;
    /// x3.rs:99-122 -- InvalidEncodingThresh when thresholds[k] > offset of codes[k], k = 0, 1
    pub fn new(block_len: usize, blocks_per_frame: usize, codes: [usize; 3], thresholds: [usize; 3]) -> Result<Self> {
        let p = Parameters { block_len, blocks_per_frame, codes, thresholds };
        check(unsafe { ffi::x3_params_validate(&p.c_struct()) })?;
        Ok(p)
    }
    pub(crate) fn c_struct(&self) -> ffi::x3_params {
        ffi::x3_params {
            block_len: self.block_len as u32,
            blocks_per_frame: self.blocks_per_frame as u32,
            codes: [self.codes[0] as u32, self.codes[1] as u32, self.codes[2] as u32],
            thresholds: [self.thresholds[0] as u32, self.thresholds[1] as u32, self.thresholds[2] as u32],
        }
    }
}
impl Default for Parameters {
    fn default() -> Self {
        Parameters { block_len: 20, blocks_per_frame: 500, codes: [0, 1, 3], thresholds: [3, 8, 20] }
    }
}

/// x3.rs:29-45 -- slice-backed channel: the natural GPU entry
pub struct Channel<'a> {
    pub id: u16,
    pub wav: &'a [i16],
    pub sample_rate: u32,
    pub params: Parameters,
}
impl<'a> Channel<'a> {
    pub fn new(id: u16, wav: &'a [i16], sample_rate: u32, params: Parameters) -> Self {
        Channel { id, wav, sample_rate, params }
    }
}
/// x3.rs:47-69 -- iterator-backed channel; drained into a Vec when encoded
pub struct IterChannel<I: Iterator<Item = i16>> {
    pub id: u16,
    pub wav: I,
    pub sample_rate: u32,
    pub params: Parameters,
}
impl<I: Iterator<Item = i16>> IterChannel<I> {
    pub fn new(id: u16, wav: impl IntoIterator<IntoIter = I>, sample_rate: u32, params: Parameters) -> Self {
        IterChannel { id, wav: wav.into_iter(), sample_rate, params }
    }
}
pub struct FrameHeader {
    pub source_id: u8,
    pub samples: u16,
    pub channels: u8,
    pub payload_len: usize,
    pub payload_crc: u16,
}
impl FrameHeader { pub const LENGTH: usize = 20; pub const KEY: u16 = 30771; }
pub struct Archive;
impl Archive { pub const ID: &'static [u8] = b"X3ARCHIV"; }
