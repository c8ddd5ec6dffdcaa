// TEST INFRASTRUCTURE.  Force-included (-include) into the three translation units of the reference that see the DecLib class layout — DecLib.cpp, vvdecimpl.cpp,
// vvdec.cpp — when oracle/_ref/libvvdec_swapped.so is built (Makefile.ref): `std::list<DecLibRecon> m_decLibRecon{ 2 }` (DecLib.h:70) and every use of it in
// DecLib.cpp then name the drop-in class instead.  This is the one-line change INTEGRATION.md asks a maintainer to make in DecLib.h, done from outside so that the
// reference's sources stay untouched.
// The drop-in class reads a few protected members of the reference's tool classes (Reshape's LMCS tables, SampleAdaptiveOffset's boundary derivation,
// AdaptiveLoopFilter's APS tables): an integration adds the `friend class b200glue::DecLibReconB200;` lines INTEGRATION.md lists; here, from outside, the access
// specifiers are opened while the reference's headers are read (as ref_shim.cpp and the reference's own unit test do).
#pragma once
#include <functional>
#include <algorithm>
#include <iostream>
#include <sstream>
#include <fstream>
#include <list>
#include <map>
#include <array>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <thread>
#include <memory>
#include <vector>
#include <string>
#include <chrono>
#include <random>
#include <unordered_map>
#include <unordered_set>
#include <set>
#include <deque>
#include <queue>
#include <exception>
#include <future>
#include <cstring>
#include <cstdarg>
#include <cmath>
#include <limits>
#include <numeric>
#include <iomanip>
#include <type_traits>
#include <tuple>
#include <utility>
#include <bitset>
#include <cassert>
#define private public
#define protected public
#include "DecoderLib/DecLibRecon.h"
#include "../vvdec_b200/vvdec_glue/DecLibReconB200.h"
#undef private
#undef protected
#define DecLibRecon ::b200glue::DecLibReconB200
