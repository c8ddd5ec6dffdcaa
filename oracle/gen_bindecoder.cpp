// TEST INFRASTRUCTURE (not shipped, not on the product path).
// A second build of the unmodified reference (oracle/_ref/libvvdec_gen.so, Makefile.ref) in which ONE translation unit is replaced: this file defines the
// member functions the reference declares for `vvdec::BinDecoder` (DecoderLib/BinDecoder.h:54-92) — but instead of reading arithmetic-coded bins from the
// slice data it DRAWS every bin from a seeded generator (per-context probabilities settable from outside) and records it.  The reference's own CABACReader
// (CABACReader.cpp) then walks its syntax with those bins, so the recorded sequence is by construction a syntactically valid slice_data() for the
// parameter sets / slice header it was drawn under: every context selection, binarisation and inference is the reference's.  The recorded (context, bin)
// sequences are arithmetic-ENCODED afterwards (ref_stream.h: ref_cabac_encode) and spliced behind the slice headers, which gives a real VVC bitstream that
// the stock library (and the drop-in build) parse back to the same picture — oracle/vvc_stream.py drives the three steps.
// Nothing of the reference is copied: the class is the reference's declaration, the bodies below are ours.
#include <mutex>
#include <vector>
#include <cstdint>
#include <cstdio>
#include "DecoderLib/BinDecoder.h"
#include "CommonLib/BitStream.h"
#include "ref_stream.h"

namespace gen
{
struct Segment { int qp, initId; std::vector<int16_t> ctx; std::vector<uint8_t> bins; };
static std::mutex              g_m;
static std::vector<Segment*>   g_segs;                   // in the order of the context initialisations (= slices in decoding order with one parse thread)
static uint32_t                g_seed = 1;
static uint8_t                 g_bias[1024];             // P( bin == 1 ) * 256 per context; [1023]: bypass bins
static bool                    g_biasInit = false;
static uint32_t                g_maxEpRun = 12;          // bypass policy: after this many consecutive 1s the next bypass bin is 0 (bounds unary / Golomb prefixes)

static void initBias() { if( !g_biasInit ) { for( auto& b : g_bias ) b = 128; g_biasInit = true; } }
static inline uint32_t next( uint32_t& s ) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
}

extern "C" void gen_reset( uint32_t seed )
{
  std::lock_guard<std::mutex> l( gen::g_m );
  for( auto s : gen::g_segs ) delete s;
  gen::g_segs.clear(); gen::g_seed = seed ? seed : 1; gen::initBias();
}
// probability of a 1 (p256 / 256) for contexts [first, first + count); first == -1: the bypass bins
extern "C" void gen_set_bias( int first, int count, int p256 )
{
  gen::initBias();
  if( first < 0 ) { gen::g_bias[1023] = (uint8_t) p256; return; }
  for( int i = first; i < first + count && i < 1023; i++ ) gen::g_bias[i] = (uint8_t) p256;
}
extern "C" void gen_set_max_bypass_run( int n ) { gen::g_maxEpRun = n; }
extern "C" int  gen_num_segments() { std::lock_guard<std::mutex> l( gen::g_m ); return (int) gen::g_segs.size(); }
// info[0..2] = qp, initId, bins; the arrays may be null to query the size
extern "C" long gen_segment( int i, int* info, int16_t* ctx, uint8_t* bins, long cap )
{
  std::lock_guard<std::mutex> l( gen::g_m );
  if( i < 0 || i >= (int) gen::g_segs.size() ) return -1;
  const gen::Segment& s = *gen::g_segs[i];
  if( info ) { info[0] = s.qp; info[1] = s.initId; info[2] = (int) s.ctx.size(); }
  if( ctx && bins )
  {
    if( (long) s.ctx.size() > cap ) return -2;
    memcpy( ctx, s.ctx.data(), s.ctx.size() * sizeof( int16_t ) ); memcpy( bins, s.bins.data(), s.bins.size() );
  }
  return (long) s.ctx.size();
}

namespace vvdec
{
// state kept in the reference's members: m_Value = index of the segment being drawn, m_Range = generator state, m_bitsNeeded = run of bypass 1s
void BinDecoder::init( InputBitstream* bitstream ) { m_Bitstream = bitstream; }
void BinDecoder::uninit()                          { m_Bitstream = nullptr; }
void BinDecoder::start()                           {}
void BinDecoder::finish()                          { while( m_Bitstream && m_Bitstream->getNumBitsLeft() >= 8 ) m_Bitstream->readByte(); }   // the dummy payload
void BinDecoder::align()                           {}

void BinDecoder::reset( int qp, int initId )
{
  m_Ctx.init( qp, initId );
  std::lock_guard<std::mutex> l( gen::g_m );
  gen::initBias();
  gen::g_segs.push_back( new gen::Segment{ qp, initId, {}, {} } );
  m_Value = uint32_t( gen::g_segs.size() - 1 );
  m_Range = gen::g_seed * 2654435761u + m_Value * 40503u + 1u; if( !m_Range ) m_Range = 1;
  m_bitsNeeded = 0;
}

static inline unsigned draw( uint32_t& state, unsigned p256 ) { return ( gen::next( state ) >> 8 & 0xff ) < p256; }
static inline void     record( uint32_t seg, int ctx, unsigned bin ) { gen::Segment& s = *gen::g_segs[seg]; s.ctx.push_back( (int16_t) ctx ); s.bins.push_back( (uint8_t) bin ); }

unsigned BinDecoder::decodeBin( unsigned ctxId )
{
  const unsigned bin = draw( m_Range, gen::g_bias[ctxId < 1023 ? ctxId : 1022] );
  record( m_Value, (int) ctxId, bin );
  return bin;
}
unsigned BinDecoder::decodeBinEP()
{
  unsigned bin = draw( m_Range, gen::g_bias[1023] );
  if( bin && (uint32_t) m_bitsNeeded >= gen::g_maxEpRun ) bin = 0;
  m_bitsNeeded = bin ? m_bitsNeeded + 1 : 0;
  record( m_Value, -1, bin );
  return bin;
}
unsigned BinDecoder::decodeBinsEP( unsigned numBins )
{
  unsigned v = 0;
  for( unsigned i = 0; i < numBins; i++ ) { const unsigned bin = draw( m_Range, 128 ); record( m_Value, -1, bin ); v = ( v << 1 ) | bin; }   // fixed-length fields: uniform
  m_bitsNeeded = 0;
  return v;
}
unsigned BinDecoder::decodeAlignedBinsEP( unsigned numBins ) { return decodeBinsEP( numBins ); }

// the binarisation of abs_remainder / dec_abs_level (9.3.3.11-12) as the reference reads it (BinDecoder.cpp:191-218), on drawn bins
unsigned BinDecoder::decodeRemAbsEP( unsigned goRicePar, unsigned cutoff, int maxLog2TrDynamicRange )
{
  const unsigned maxPrefix = 32 - maxLog2TrDynamicRange;
  unsigned prefix = 0;
  m_bitsNeeded = 0;
  while( prefix < maxPrefix && decodeBinEP() ) prefix++;
  unsigned length = goRicePar, offset;
  if( prefix < cutoff ) offset = prefix << goRicePar;
  else
  {
    offset  = ( ( 1u << ( prefix - cutoff ) ) + cutoff - 1 ) << goRicePar;
    length += prefix == maxPrefix ? maxLog2TrDynamicRange - goRicePar : prefix - cutoff;
  }
  return offset + decodeBinsEP( length );
}
unsigned BinDecoder::decodeBinTrm()
{
  record( m_Value, -2, 1 );                          // the reference only reads the terminating bin where the (sub)stream has to end (DecSlice.cpp:162-176)
  return 1;
}
#if ENABLE_TRACING
unsigned BinDecoder::getNumBitsRead() const { return 0; }
#endif
}   // namespace vvdec
