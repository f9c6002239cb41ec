//! X3Error with the reference's variants (src/error.rs:27-62), built from the C-ABI return codes.
#[derive(Debug)]
pub enum X3Error {
    Io(std::io::Error),
    Hound(hound::Error),
    InvalidEncodingThresh,
    OutOfBoundsInverse,
    MoreThanOneChannel,
    ArchiveHeaderXMLInvalid,
    ArchiveHeaderXMLRiceCode,
    ArchiveHeaderXMLInvalidKey,
    FrameLength,
    FrameHeaderInvalidKey,
    FrameHeaderInvalidPayloadLen,
    FrameHeaderInvalidHeaderCRC,
    FrameHeaderInvalidPayloadCRC,
    FrameDecodeInvalidFType,
    FrameDecodeInvalidBPF,
    FrameDecodeUnexpectedEnd,
    ByteWriterInsufficientMemory,
    /// not in the reference: CUDA failure, unsupported parameters, invalid argument (carries the code)
    Backend(i32),
}
pub type Result<T> = core::result::Result<T, X3Error>;

pub(crate) fn check(code: i32) -> Result<()> {
    use X3Error::*;
    Err(match code {
        0 => return Ok(()),
        -1 => InvalidEncodingThresh,
        -2 => OutOfBoundsInverse,
        -3 => MoreThanOneChannel,
        -4 => ArchiveHeaderXMLInvalid,
        -5 => ArchiveHeaderXMLRiceCode,
        -6 => ArchiveHeaderXMLInvalidKey,
        -7 => FrameLength,
        -8 => FrameHeaderInvalidKey,
        -9 => FrameHeaderInvalidPayloadLen,
        -10 => FrameHeaderInvalidHeaderCRC,
        -11 => FrameHeaderInvalidPayloadCRC,
        -12 => FrameDecodeInvalidFType,
        -13 => FrameDecodeInvalidBPF,
        -14 => FrameDecodeUnexpectedEnd,
        -15 => ByteWriterInsufficientMemory,
        -16 => Io(std::io::Error::from(std::io::ErrorKind::UnexpectedEof)),
        c => Backend(c),
    })
}
impl From<std::io::Error> for X3Error { fn from(e: std::io::Error) -> Self { X3Error::Io(e) } }
impl From<hound::Error> for X3Error { fn from(e: hound::Error) -> Self { X3Error::Hound(e) } }
