// albatross_b200 — umbrella header of the C++ trait layer (the host side of the drop-in boundary).
//
//   #include <albatross_b200/albatross.hpp>          // instead of <albatross/GP>, <albatross/SparseGP>
//   namespace albatross = albatross_b200;            // or -DALBATROSS_B200_AS_ALBATROSS
//
// and link libalbatross_b200.so.  The layer is header-only C++17 over include/albatross_b200.h; it
// re-creates the reference's template concepts for the exact-GP hot path with the
// same names and signatures.  Types without a device form fail to compile; nothing falls back to
// the host.
#pragma once

#include "core.hpp"
#include "covariance.hpp"
#include "device.hpp"
#include "gp.hpp"
#include "linalg.hpp"
#include "linalg_types.hpp"
#include "parameters.hpp"
#include "sparse_gp.hpp"
#include "tune.hpp"

#ifdef ALBATROSS_B200_AS_ALBATROSS
namespace albatross = albatross_b200;
#endif
