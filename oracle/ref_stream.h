// TEST INFRASTRUCTURE (not shipped, not on the product path): bitstream-level access to the compiled reference.
//   ref_decode_stream()  — the reference's public decoder (vvdec_decoder_open / vvdec_decode / vvdec_flush, include/vvdec/vvdec.h.in:548-620) over an
//                          Annex-B byte string, frames copied out as 16-bit planes.
//   ref_cabac_encode()   — an arithmetic ENCODER for recorded (context, bin) sequences: the standard's 9.3.4.x coder written from the decoder's view
//                          (BinDecoder.cpp:81-340: range 510, 9-bit range / 16+ bit value registers, terminate at range - 2) with the reference's
//                          own probability models (Contexts.h: BinProbModel::lpsmps / update / getRenormBitsLPS, Ctx::init( qp, initId )), so the
//                          context states evolve exactly as they do in the reference when it reads the bins back.
//   ref_ctx_range()      — offset and size of a named context set (Contexts.h ContextSetCfg).
// Included by ref_shim.cpp (stock library) and gen_bindecoder.cpp (library whose BinDecoder draws the bins).
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>
#include <string>
#include "vvdec/vvdec.h"
#include "CommonLib/Contexts.h"

namespace refstream
{
using namespace vvdec;

// ---- arithmetic encoder -------------------------------------------------------------------------------------------------------------------------------
// Register layout of the HM / VTM encoder (low: 10 fraction bits above the 9-bit range window, carries resolved through a buffered byte + 0xff run).
struct BinEncoder
{
  std::vector<uint8_t>& out;
  uint32_t low = 0, range = 510;
  int      bitsLeft = 23, numBufferedBytes = 0;
  uint8_t  bufferedByte = 0xff;
  explicit BinEncoder( std::vector<uint8_t>& o ) : out( o ) {}

  void writeOut()
  {
    const unsigned leadByte = low >> ( 24 - bitsLeft );
    bitsLeft += 8;
    low &= 0xffffffffu >> bitsLeft;
    if( leadByte == 0xff ) { numBufferedBytes++; return; }
    if( numBufferedBytes > 0 )
    {
      const unsigned carry = leadByte >> 8;
      out.push_back( uint8_t( bufferedByte + carry ) );
      const uint8_t fill = uint8_t( 0xff + carry );
      for( ; numBufferedBytes > 1; numBufferedBytes-- ) out.push_back( fill );
      bufferedByte = uint8_t( leadByte & 0xff );
    }
    else { numBufferedBytes = 1; bufferedByte = uint8_t( leadByte ); }
  }
  void encodeBin( BinProbModel& m, unsigned bin )
  {
    unsigned lps, mps; m.lpsmps( range, lps, mps );
    range -= lps;
    if( bin != mps )
    {
      const int n = m.getRenormBitsLPS( lps );
      bitsLeft -= n; low = ( low + range ) << n; range = lps << n;
    }
    else if( range < 256 )
    {
      const int n = m.getRenormBitsRange( range );
      bitsLeft -= n; low <<= n; range <<= n;
    }
    m.update( bin );
    if( bitsLeft < 12 ) writeOut();
  }
  void encodeBinEP( unsigned bin )
  {
    low <<= 1;
    if( bin ) low += range;
    if( --bitsLeft < 12 ) writeOut();
  }
  void encodeBinTrm( unsigned bin )
  {
    range -= 2;
    if( bin ) { low += range; low <<= 7; range = 2 << 7; bitsLeft -= 7; }
    else if( range < 256 ) { low <<= 1; range <<= 1; bitsLeft--; }
    if( bitsLeft < 12 ) writeOut();
  }
  // after the terminating bin: flush the registers, then the stop bit and the alignment zeros (rbsp_slice_trailing_bits)
  void finish()
  {
    uint64_t bits = 0; int nbits = 0;                                        // bit accumulator for the tail
    auto put = [&]( uint32_t v, int n ) { for( int i = n - 1; i >= 0; i-- ) { bits = ( bits << 1 ) | ( ( v >> i ) & 1 ); if( ++nbits == 8 ) { out.push_back( uint8_t( bits ) ); bits = 0; nbits = 0; } } };
    if( low >> ( 32 - bitsLeft ) )
    {
      out.push_back( uint8_t( bufferedByte + 1 ) );
      for( ; numBufferedBytes > 1; numBufferedBytes-- ) out.push_back( 0x00 );
      low -= 1u << ( 32 - bitsLeft );
    }
    else
    {
      if( numBufferedBytes > 0 ) out.push_back( bufferedByte );
      for( ; numBufferedBytes > 1; numBufferedBytes-- ) out.push_back( 0xff );
    }
    put( low >> 8, 24 - bitsLeft );
    put( 1, 1 );
    while( nbits ) put( 0, 1 );
  }
};
inline int& lastHashErrors() { static int n = 0; return n; }
}   // namespace refstream

// pictures of the last ref_decode_stream() call whose decoded-picture-hash SEI did not match what the decoder reconstructed (vvdec_get_hash_error_count)
extern "C" int ref_last_hash_errors() { return refstream::lastHashErrors(); }

// One CABAC segment (what lies between two context initialisations: a slice without tiles / wavefronts).  ctx[i] >= 0: context-coded bin with that
// context index; -1: bypass bin; -2: terminating bin (the last entry of a segment, value 1).  Returns the bytes written (the slice_data() bytes,
// rbsp_slice_trailing_bits included) or -1 if `cap` is too small.
extern "C" long ref_cabac_encode( int qp, int initId, const int16_t* ctx, const uint8_t* bins, long n, uint8_t* outBuf, long cap )
{
  using namespace refstream;
  vvdec::Ctx models; models.init( qp, initId );
  std::vector<uint8_t> out; out.reserve( (size_t) n / 4 + 16 );
  BinEncoder enc( out );
  for( long i = 0; i < n; i++ )
  {
    if( ctx[i] >= 0 )       enc.encodeBin( models[ctx[i]], bins[i] );
    else if( ctx[i] == -1 ) enc.encodeBinEP( bins[i] );
    else                    enc.encodeBinTrm( bins[i] );
  }
  enc.finish();
  if( (long) out.size() > cap ) return -1;
  memcpy( outBuf, out.data(), out.size() );
  return (long) out.size();
}

#define REFSTREAM_CTX_SETS( X ) \
  X( SplitFlag ) X( SplitQtFlag ) X( SplitHvFlag ) X( Split12Flag ) X( ModeConsFlag ) X( SkipFlag ) X( MergeFlag ) X( RegularMergeFlag ) X( MergeIdx ) \
  X( PredMode ) X( MultiRefLineIdx ) X( IntraLumaPlanarFlag ) X( CclmModeFlag ) X( CclmModeIdx ) X( MipFlag ) X( DeltaQP ) X( InterDir ) X( RefPic ) \
  X( MmvdFlag ) X( MmvdMergeIdx ) X( MmvdStepMvpIdx ) X( SubblockMergeFlag ) X( AffineFlag ) X( AffineType ) X( AffMergeIdx ) X( Mvd ) X( BDPCMMode ) \
  X( QtRootCbf ) X( ACTFlag ) X( TsSigCoeffGroup ) X( TsSigFlag ) X( TsParFlag ) X( TsGtxFlag ) X( TsLrg1Flag ) X( TsResidualSign ) X( MVPIdx ) \
  X( SaoMergeFlag ) X( SaoTypeIdx ) X( MTSIndex ) X( LFNSTIdx ) X( RdpcmFlag ) X( RdpcmDir ) X( SbtFlag ) X( SbtQuadFlag ) X( SbtHorFlag ) \
  X( SbtPosFlag ) X( ChromaQpAdjFlag ) X( ChromaQpAdjIdc ) X( ImvFlag ) X( CcAlfFilterControlFlag ) X( BcwIdx ) X( ctbAlfFlag ) X( ctbAlfAlternative ) X( AlfUseTemporalFilt ) \
  X( CiipFlag ) X( SmvdFlag ) X( IBCFlag ) X( ISPMode ) X( JointCbCrFlag )
#define REFSTREAM_CTX_ARRAYS( X ) X( IPredMode, 2 ) X( QtCbf, 3 ) X( SigCoeffGroup, 2 ) X( LastX, 2 ) X( LastY, 2 ) X( SigFlag, 6 ) X( ParFlag, 2 ) X( GtxFlag, 4 )

// offset of a context set by name ("SplitFlag", "QtCbf1" for an element of an array of sets); -1 if unknown.  "total": the number of contexts (a set
// ends where the next one starts).
extern "C" int ref_ctx_offset( const char* name )
{
  using namespace vvdec;
#define X( n ) if( !strcmp( name, #n ) ) return Ctx::n();
  REFSTREAM_CTX_SETS( X )
#undef X
#define X( n, k ) if( !strncmp( name, #n, sizeof( #n ) - 1 ) && name[sizeof( #n ) - 1] >= '0' && name[sizeof( #n ) - 1] < '0' + k && !name[sizeof( #n )] ) return Ctx::n[name[sizeof( #n ) - 1] - '0']();
  REFSTREAM_CTX_ARRAYS( X )
#undef X
  if( !strcmp( name, "total" ) ) return (int) ContextSetCfg::NumberOfContexts;
  return -1;
}
// all names ref_ctx_offset() knows, '\n'-separated
extern "C" const char* ref_ctx_names()
{
#define X( n ) #n "\n"
#define Y( n, k ) #n ":" #k "\n"
  return REFSTREAM_CTX_SETS( X ) REFSTREAM_CTX_ARRAYS( Y );
#undef X
#undef Y
}

// The whole stream through the reference's public API.  `stream`: Annex-B access units, auOffsets[0..nAu] their byte ranges (each vvdec_decode call takes
// one).  Frames come out in output order as 16-bit planes, appended to `out` (Y, Cb, Cr per frame, tightly packed); dims[0..5] = luma w, h, chroma w, h,
// bit depth, frames.  frameDims (may be null, 4 ints per frame, room for maxFrames): the plane sizes of every frame (streams that change resolution).
// Returns the number of frames, or a negative vvdec error code (message in errBuf).
extern "C" int ref_decode_stream2( const uint8_t* stream, const long* auOffsets, int nAu, int threads, int16_t* out, long outCap, int* dims, char* errBuf, int errCap, int* frameDims, int maxFrames )
{
  vvdecParams params; vvdec_params_default( &params );
  params.threads = threads; params.logLevel = VVDEC_SILENT; params.errHandlingFlags = VVDEC_ERR_HANDLING_OFF;
  params.verifyPictureHash = true;                                           // decoded-picture-hash SEIs are checked by the decoder itself (ref_last_hash_errors)
  if( threads <= 1 ) params.parseDelay = 0;
  vvdecDecoder* dec = vvdec_decoder_open( &params );
  if( !dec ) { if( errBuf ) snprintf( errBuf, errCap, "vvdec_decoder_open failed" ); return -1000; }
  int  frames = 0, rc = 0; long used = 0;
  auto fail = [&]( int code ) { if( errBuf ) snprintf( errBuf, errCap, "%s | %s", vvdec_get_last_error( dec ), vvdec_get_last_additional_error( dec ) ); rc = code < 0 ? code : -999; };
  auto take = [&]( vvdecFrame* f )
  {
    if( !f ) return;
    for( unsigned c = 0; c < f->numPlanes; c++ )
    {
      const vvdecPlane& p = f->planes[c];
      if( c < 2 ) { dims[2 * c] = p.width; dims[2 * c + 1] = p.height; if( frameDims && frames < maxFrames ) { frameDims[4 * frames + 2 * c] = p.width; frameDims[4 * frames + 2 * c + 1] = p.height; } }
      if( used + (long) p.width * p.height > outCap ) { rc = -998; break; }
      for( unsigned y = 0; y < p.height; y++ )
      {
        const unsigned char* row = p.ptr + (size_t) y * p.stride;
        int16_t* dst = out + used + (size_t) y * p.width;
        if( p.bytesPerSample == 2 ) memcpy( dst, row, p.width * 2 ); else for( unsigned x = 0; x < p.width; x++ ) dst[x] = row[x];
      }
      used += (long) p.width * p.height;
    }
    if( f->numPlanes == 1 && frameDims && frames < maxFrames ) frameDims[4 * frames + 2] = frameDims[4 * frames + 3] = 0;
    dims[4] = f->bitDepth; frames++;
    vvdec_frame_unref( dec, f );
  };
  vvdecAccessUnit* au = vvdec_accessUnit_alloc();
  for( int i = 0; i < nAu && rc == 0; i++ )
  {
    const int len = int( auOffsets[i + 1] - auOffsets[i] );
    vvdec_accessUnit_alloc_payload( au, len + 16 );
    memcpy( au->payload, stream + auOffsets[i], len ); au->payloadUsedSize = len;
    au->cts = i; au->ctsValid = true; au->dts = i; au->dtsValid = true;
    vvdecFrame* f = nullptr;
    int r = vvdec_decode( dec, au, &f );
    vvdec_accessUnit_free_payload( au ); au->payload = nullptr;              // (the call leaves the dangling pointer, which vvdec_accessUnit_free() would free again)
    if( r != VVDEC_OK && r != VVDEC_TRY_AGAIN ) { fail( r ); break; }
    take( f );
  }
  while( rc == 0 )
  {
    vvdecFrame* f = nullptr;
    int r = vvdec_flush( dec, &f );
    if( r != VVDEC_OK && r != VVDEC_EOF ) { fail( r ); break; }
    take( f );
    if( r == VVDEC_EOF || !f ) break;
  }
  vvdec_accessUnit_free( au );
  refstream::lastHashErrors() = vvdec_get_hash_error_count( dec );
  vvdec_decoder_close( dec );
  dims[5] = frames;
  return rc ? rc : frames;
}
extern "C" int ref_decode_stream( const uint8_t* stream, const long* auOffsets, int nAu, int threads, int16_t* out, long outCap, int* dims, char* errBuf, int errCap )
{
  return ref_decode_stream2( stream, auOffsets, nAu, threads, out, outCap, dims, errBuf, errCap, nullptr, 0 );
}
