// flatten_output.h — host-side glue of the output path: FilmGrain state -> b200_film_grain.  Lives INSIDE a VVdeC build; no pixel arithmetic.
// FilmGrain keeps m_impl / m_line_seeds / fgs private and FilmGrainImpl its tables protected (FilmGrain.h:87-95, FilmGrainImpl.h:96-106):
// a maintainer adds `friend struct b200glue::FilmGrainTables;` to both classes; oracle/ref_shim.cpp (test code) compiles with access opened.
// Pinned by tests/test_film_grain_oracle_vs_ref.py.
#pragma once
#include <vector>
#include <string.h>
#include "vvdec_b200.h"
#include "FilmGrain/FilmGrain.h"
#include "FilmGrain/FilmGrainImpl.h"

namespace b200glue
{
using namespace vvdec;

struct FilmGrainTables
{
  std::vector<int8_t>   pattern;
  std::vector<uint32_t> lineSeeds;
  uint8_t               sLUT[3 * 256], pLUT[3 * 256];
  b200_film_grain       fg;

  // call after FilmGrain::setDepth / setColorFormat / prepareBlockSeeds of the frame (VVDecImpl::xAddGrain, vvdecimpl.cpp:905-907)
  void flatten( const FilmGrain& f )
  {
    const FilmGrainImpl& im = *f.m_impl;
    pattern.resize( 2 * 8 * 64 * 64 );
    for( int k = 0; k < 2; k++ ) for( int i = 0; i < 8; i++ ) memcpy( &pattern[( k * 8 + i ) * 4096], im.pattern[k][i], 4096 );   // [8] only serves interpolation (off)
    memcpy( sLUT, im.sLUT, sizeof( sLUT ) );
    memcpy( pLUT, im.pLUT, sizeof( pLUT ) );
    lineSeeds = f.m_line_seeds;
    fg.pattern = pattern.data(); fg.sLUT = sLUT; fg.pLUT = pLUT; fg.lineSeeds = lineSeeds.data();
    fg.scaleShift = im.scale_shift;
    for( int c = 0; c < 3; c++ ) fg.compPresent[c] = f.fgs.comp_model_present_flag[c];
  }
};

}   // namespace b200glue
