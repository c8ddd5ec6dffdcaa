// TEST INFRASTRUCTURE (not shipped): entry points of oracle/_ref/libvvdec_swapped.so — the unmodified reference with b200glue::DecLibReconB200 compiled in behind
// the DecLibRecon seam (swap_recon.h).  ref_decode_stream() (ref_stream.h) decodes a bitstream through the reference's public API; swapped_set_hooks() selects how
// the drop-in class gets its results:
//   dryRun = 0: the product path — libvvdec_b200.so on the GPU (the GPU tests);
//   dryRun = 1: no device; where the device result would be fetched the callback receives the flattened work lists and fills planes / DMVR deltas — the CPU tests
//               put the oracle chain there, so the whole decoder runs bitstream -> parser -> glue host stages -> oracle on a machine without a GPU.
#define B200_GLUE_TEST_HOOKS 1
#include "swap_recon.h"
#include "ref_stream.h"

typedef void ( *swapped_picture_fn )( void* user, const b200_picture* lists, const b200_geom* geom, int32_t* dmvrDeltas, size_t numDmvr, int16_t* const planes[3], const ptrdiff_t strides[3], int poc );

typedef void ( *swapped_load_fn )( void* user, int slot, const int16_t* const planes[3], const ptrdiff_t strides[3], const b200_geom* geom );

extern "C" void swapped_set_hooks( int dryRun, swapped_picture_fn fn, swapped_load_fn load, void* user )
{
  auto& h = b200glue::DecLibReconB200::testHooks();
  h.dryRun = dryRun != 0; h.picture = fn; h.loadSlot = load; h.user = user;
}

// instances created from now on complete their pictures in a pool task (DecLibReconB200::setAsyncFinish) instead of in waitForPrevDecompressedPic()
extern "C" void swapped_set_async_finish( int on ) { b200glue::DecLibReconB200::asyncFinishDefault() = on != 0; }
