//! x3-b200: the x3 crate's public API (x3::Parameters, x3::Channel, encoder::encode, decoder::decode_frame,
//! encodefile::wav_to_x3a, decodefile::x3a_to_wav) over hand-written sm_100a CUDA.  SOURCE ONLY -- see Cargo.toml.
pub mod error;
pub mod ffi;
pub mod x3;

pub mod bytewriter {
    //! src/bytewriter.rs: the trait is the reference's, item for item (bytewriter.rs:14-22), so that a caller's own
    //! `ByteWriter` plugs in unchanged: `align<const N>`, `write_all(impl AsRef<[u8]>)`, `flush`, `seek`,
    //! `stream_position`.  The GPU path knows every frame's size before it writes, so it never seeks itself; `seek` is
    //! there for callers (and for the trait to be the same).
    use crate::error::{Result, X3Error};
    #[cfg(feature = "std")]
    pub use std::io::SeekFrom;
    #[cfg(not(feature = "std"))]
    pub enum SeekFrom { Start(u64), End(i64), Current(i64) }

    pub trait ByteWriter {
        fn align<const N: usize>(&mut self) -> Result<usize>;
        fn write_all(&mut self, value: impl AsRef<[u8]>) -> Result<()>;
        fn flush(&mut self) -> Result<()>;
        fn seek(&mut self, pos: SeekFrom) -> Result<u64>;
        fn stream_position(&mut self) -> Result<u64>;
    }

    /// A caller-owned slice; never grows (ByteWriterInsufficientMemory, bytewriter.rs:72-74,88-90).
    pub struct SliceByteWriter<'a> { slice: &'a mut [u8], at: usize, high_water: usize }
    impl<'a> SliceByteWriter<'a> {
        pub fn new(slice: &'a mut [u8]) -> Self { SliceByteWriter { slice, at: 0, high_water: 0 } }
    }
    impl<'a> ByteWriter for SliceByteWriter<'a> {
        fn align<const N: usize>(&mut self) -> Result<usize> {
            let pad = (N - self.at % N) % N;
            if pad > 0 { self.write_all(&[0u8; N][..pad])?; }
            Ok(pad)
        }
        fn write_all(&mut self, value: impl AsRef<[u8]>) -> Result<()> {
            let v = value.as_ref();
            let end = self.at.checked_add(v.len()).ok_or(X3Error::ByteWriterInsufficientMemory)?;
            if end > self.slice.len() { return Err(X3Error::ByteWriterInsufficientMemory); }
            self.slice[self.at..end].copy_from_slice(v);
            self.at = end;
            if end > self.high_water { self.high_water = end; }
            Ok(())
        }
        fn flush(&mut self) -> Result<()> { Ok(()) }
        fn seek(&mut self, pos: SeekFrom) -> Result<u64> {
            let target: i128 = match pos {
                SeekFrom::Start(p) => p as i128,
                SeekFrom::End(d) => self.high_water as i128 + d as i128,
                SeekFrom::Current(d) => self.at as i128 + d as i128,
            };
            if target < 0 || target > self.slice.len() as i128 { return Err(X3Error::ByteWriterInsufficientMemory); }
            self.at = target as usize;
            if self.at > self.high_water { self.high_water = self.at; }
            Ok(self.at as u64)
        }
        fn stream_position(&mut self) -> Result<u64> { Ok(self.at as u64) }
    }

    /// Any `Write + Seek` (a file behind a BufWriter in wav_to_x3a).
    #[cfg(feature = "std")]
    pub struct StreamByteWriter<'a, W: std::io::Write + std::io::Seek> { writer: &'a mut W }
    #[cfg(feature = "std")]
    impl<'a, W: std::io::Write + std::io::Seek> StreamByteWriter<'a, W> {
        pub fn new(writer: &'a mut W) -> Self { StreamByteWriter { writer } }
    }
    #[cfg(feature = "std")]
    impl<'a, W: std::io::Write + std::io::Seek> ByteWriter for StreamByteWriter<'a, W> {
        fn align<const N: usize>(&mut self) -> Result<usize> {
            let pad = (N - (self.writer.stream_position()? as usize) % N) % N;
            if pad > 0 { self.write_all(&[0u8; N][..pad])?; }
            Ok(pad)
        }
        fn write_all(&mut self, value: impl AsRef<[u8]>) -> Result<()> { Ok(std::io::Write::write_all(self.writer, value.as_ref())?) }
        fn flush(&mut self) -> Result<()> { Ok(self.writer.flush()?) }
        fn seek(&mut self, pos: SeekFrom) -> Result<u64> { Ok(self.writer.seek(pos)?) }
        fn stream_position(&mut self) -> Result<u64> { Ok(self.writer.stream_position()?) }
    }
}

pub mod encoder {
    //! src/encoder.rs over x3_encode_host
    use crate::bytewriter::ByteWriter;
    use crate::error::{check, Result, X3Error};
    use crate::{ffi, x3};

    fn encode_slice(wav: &[i16], params: &x3::Parameters, stats: &mut [usize; 6]) -> Result<Vec<u8>> {
        let p = params.c_struct();
        let cap = unsafe { ffi::x3_encode_bound(wav.len(), &p) };
        let mut out = vec![0u8; cap];
        let (mut len, mut st) = (0usize, ffi::x3_stats::default());
        check(unsafe { ffi::x3_encode_host(wav.as_ptr(), wav.len(), &p, out.as_mut_ptr(), cap, &mut len, &mut st) })?;
        out.truncate(len);
        for k in 0..6 { stats[k] += st.samples_by_mode[k] as usize; }
        Ok(out)
    }
    fn print_stats(stats: &[usize; 6]) {
        let t = stats.iter().sum::<usize>() as f32;  // encoder.rs:96-108
        println!("\nStatistics:\n  Rice-0: {:.4}%\n  Rice-1: {:.4}%\n  Rice-2: {:.4}%\n  Rice-3: {:.4}%\n  BFP: {:.4}%\n  Pass-through {:.4}%\n",
                 stats[0] as f32 / t * 100.0, stats[1] as f32 / t * 100.0, stats[2] as f32 / t * 100.0,
                 stats[3] as f32 / t * 100.0, stats[4] as f32 / t * 100.0, stats[5] as f32 / t * 100.0);
    }
    /// encoder::encode (encoder.rs:51): current signature, IterChannel + ByteWriter
    pub fn encode<I: Iterator<Item = i16>, W: ByteWriter>(channels: &mut [&mut x3::IterChannel<I>], writer: &mut W) -> Result<()> {
        if channels.len() > 1 { return Err(X3Error::MoreThanOneChannel); }
        let ch = &mut channels[0];
        let wav: Vec<i16> = ch.wav.by_ref().collect();
        encode_channel(&x3::Channel::new(ch.id, &wav, ch.sample_rate, x3::Parameters { ..x3::Parameters::new(ch.params.block_len, ch.params.blocks_per_frame, ch.params.codes, ch.params.thresholds)? }), writer)
    }
    /// README form (README.md:43-50): slice-backed channel
    pub fn encode_channel<W: ByteWriter>(ch: &x3::Channel, writer: &mut W) -> Result<()> {
        let mut stats = [0usize; 6];
        if !ch.wav.is_empty() {
            writer.align::<2>()?;  // encoder.rs:182
            let data = encode_slice(ch.wav, &ch.params, &mut stats)?;
            writer.write_all(&data)?;
        }
        print_stats(&stats);
        Ok(())
    }
    /// encoder::encode_frame (encoder.rs:175)
    pub fn encode_frame<W: ByteWriter>(wav: &[i16], writer: &mut W, params: &x3::Parameters, stats: &mut [usize; 6]) -> Result<()> {
        let p = params.c_struct();
        let cap = unsafe { ffi::x3_encode_frame_bound(wav.len(), &p) }.max(32);   // 2.75 bytes per sample at block_len 1
        let mut out = vec![0u8; cap];
        let (mut len, mut st) = (0usize, ffi::x3_stats::default());
        writer.align::<2>()?;
        check(unsafe { ffi::x3_encode_frame_host(wav.as_ptr(), wav.len(), &p, out.as_mut_ptr(), cap, &mut len, &mut st) })?;
        for k in 0..6 { stats[k] += st.samples_by_mode[k] as usize; }
        writer.write_all(&out[..len])
    }
    /// encoder::write_frame_header (encoder.rs:122)
    pub fn write_frame_header(num_samples: usize, id: u8, payload_len: usize, payload_crc: u16) -> [u8; 20] {
        let mut h = [0u8; 20];
        unsafe { ffi::x3_write_frame_header(num_samples, id, payload_len, payload_crc, h.as_mut_ptr()) };
        h
    }
}

pub mod decoder {
    //! src/decoder.rs over x3_decode_frame_host / x3_read_frame_header
    use crate::error::{check, Result};
    use crate::{ffi, x3};
    pub fn read_frame_header(bytes: &[u8]) -> Result<x3::FrameHeader> {
        let mut h = ffi::x3_frame_header::default();
        check(unsafe { ffi::x3_read_frame_header(bytes.as_ptr(), bytes.len(), &mut h) })?;
        Ok(x3::FrameHeader { source_id: h.source_id, samples: h.samples, channels: h.channels, payload_len: h.payload_len as usize, payload_crc: h.payload_crc })
    }
    pub fn decode_frame(x3_bytes: &mut [u8], wav_buf: &mut [i16], params: &x3::Parameters, samples: usize) -> Result<Option<usize>> {
        let mut n = 0usize;
        check(unsafe { ffi::x3_decode_frame_host(x3_bytes.as_ptr(), x3_bytes.len(), &params.c_struct(), wav_buf.as_mut_ptr(), wav_buf.len(), samples, &mut n) })?;
        Ok(Some(n))
    }
    /// whole frame stream (what X3aReader's loop does), all frames before the first bad one
    pub fn decode_stream(frames: &[u8], params: &x3::Parameters, pcm: &mut [i16]) -> (i32, ffi::x3_decode_result, usize) {
        let (mut n, mut r) = (0usize, ffi::x3_decode_result::default());
        let rc = unsafe { ffi::x3_decode_host(frames.as_ptr(), frames.len(), &params.c_struct(), pcm.as_mut_ptr(), pcm.len(), &mut n, &mut r) };
        (rc, r, n)
    }
}

pub mod encodefile {
    //! src/encodefile.rs:48-138
    use crate::bytewriter::{ByteWriter, StreamByteWriter};
    use crate::error::Result;
    use crate::{encoder, ffi, x3};
    pub fn wav_to_x3a<P: AsRef<std::path::Path>>(wav_filename: P, x3a_filename: P) -> Result<()> {
        let mut reader = hound::WavReader::open(wav_filename).unwrap();
        assert_eq!(reader.spec().bits_per_sample, 16);
        assert_eq!(reader.spec().channels, 1);
        let fs = reader.spec().sample_rate;
        let wav: Vec<i16> = reader.samples::<i16>().map(|x| x.unwrap()).collect();
        let params = x3::Parameters::default();
        let xml = format!("<X3ARCH PROG=\"x3new.m\" VERSION=\"2.0\" /><CFG ID=\"0\" FTYPE=\"XML\" /><CFG ID=\"1\" FTYPE=\"WAV\"><FS UNIT=\"Hz\">{}</FS><SUFFIX>wav</SUFFIX><CODEC TYPE=\"X3\" VERS=\"2\"><BLKLEN>{}</BLKLEN><CODES N=\"4\">RICE{},RICE{},RICE{},BFP</CODES><FILTER>DIFF</FILTER><NBITS>16</NBITS><T N=\"3\">{},{},{}</T></CODEC></CFG>",
                          fs, params.block_len, params.codes[0], params.codes[1], params.codes[2], params.thresholds[0], params.thresholds[1], params.thresholds[2]);
        let mut payload = xml.into_bytes();
        if payload.len() % 2 == 1 { payload.push(0); }
        let crc = unsafe { ffi::x3_crc16(payload.as_ptr(), payload.len()) };
        let mut file = std::io::BufWriter::new(std::fs::File::create(x3a_filename)?);
        let mut w = StreamByteWriter::new(&mut file);
        w.write_all(x3::Archive::ID)?;
        w.write_all(&encoder::write_frame_header(0, 0, payload.len(), crc))?;
        w.write_all(&payload)?;
        encoder::encode_channel(&x3::Channel::new(0, &wav, fs, params), &mut w)
    }
}

pub mod decodefile {
    //! src/decodefile.rs:189-303 -- the whole file is decoded by one GPU call
    use crate::error::{check, Result, X3Error};
    use crate::{decoder, x3};

    fn first_text<'a>(xml: &'a str, tag: &str) -> Option<&'a str> {
        let open = format!("<{}", tag);
        let mut at = 0;
        while let Some(i) = xml[at..].find(&open) {
            let s = at + i;
            let gt = s + xml[s..].find('>')?;
            let c = xml.as_bytes()[s + open.len()];
            if (c == b'>' || c == b' ') && xml.as_bytes()[gt - 1] != b'/' {
                let e = gt + 1 + xml[gt + 1..].find('<')?;
                return Some(xml[gt + 1..e].trim());
            }
            at = s + 1;
        }
        None
    }
    fn parse_xml(xml: &str) -> Result<(u32, x3::Parameters)> {
        let fs = first_text(xml, "FS").unwrap();
        let bl = first_text(xml, "BLKLEN").unwrap();
        let codes = first_text(xml, "CODES").unwrap();
        let th = first_text(xml, "T").unwrap();
        println!("sample rate: {}\nblock length: {}\nRice codes: {}\nthresholds: {}", fs, bl, codes, th);
        let mut ids = Vec::new();
        for w in codes.split(',') {
            match w { "RICE0" => ids.push(0), "RICE1" => ids.push(1), "RICE2" => ids.push(2), "RICE3" => ids.push(3), "BFP" => (),
                      _ => return Err(X3Error::ArchiveHeaderXMLRiceCode) }
        }
        let t: Vec<usize> = th.split(',').map(|s| s.parse().unwrap()).collect();
        let p = x3::Parameters::new(bl.parse().unwrap(), x3::Parameters::DEFAULT_BLOCKS_PER_FRAME, [ids[0], ids[1], ids[2]], [t[0], t[1], t[2]])?;
        Ok((fs.parse().unwrap(), p))
    }
    pub fn x3a_to_wav<P: AsRef<std::path::Path>>(x3a_filename: P, wav_filename: P) -> Result<()> {
        let d = std::fs::read(x3a_filename).unwrap();
        if d.len() < 28 { return Err(X3Error::Io(std::io::ErrorKind::UnexpectedEof.into())); }
        if &d[..8] != x3::Archive::ID { return Err(X3Error::ArchiveHeaderXMLInvalidKey); }
        let h = decoder::read_frame_header(&d[8..28])?;
        // a truncated archive header is an Io error in the reference (read_exact, decodefile.rs:163-166), not a panic
        if d.len() < 28 + h.payload_len { return Err(X3Error::Io(std::io::ErrorKind::UnexpectedEof.into())); }
        let (fs, params) = parse_xml(&String::from_utf8_lossy(&d[28..28 + h.payload_len]))?;
        let frames = &d[28 + h.payload_len..];
        // samples of the whole frames (the decode itself stops where the reference would: decodefile.rs:107-116).
        // `pos + 20 < len`, not `len - pos > 20`: a truncated last frame leaves pos beyond the end.
        let mut total = 0usize;
        let mut pos = 0usize;
        while pos + 20 < frames.len() {
            match decoder::read_frame_header(&frames[pos..pos + 20]) {
                Ok(fh) => {
                    if pos + 20 + fh.payload_len > frames.len() { break; }   // truncated: the reference ends cleanly here
                    total += fh.samples as usize;
                    pos += 20 + fh.payload_len;
                }
                Err(_) => break,
            }
        }
        let mut pcm = vec![0i16; total.max(1)];
        let (rc, _res, n) = decoder::decode_stream(frames, &params, &mut pcm[..total]);
        let spec = hound::WavSpec { channels: 1, sample_rate: fs, bits_per_sample: 16, sample_format: hound::SampleFormat::Int };
        let mut w = hound::WavWriter::create(wav_filename, spec)?;
        for s in &pcm[..n] { w.write_sample(*s)?; }
        w.finalize()?;
        check(rc)
    }
}
