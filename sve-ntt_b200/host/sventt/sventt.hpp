// SPDX-License-Identifier: Apache-2.0
// Umbrella header, drop-in for the reference's <sventt/sventt.hpp> on the 64-bit NTT hot path.
#ifndef XNTT_SVENTT_HPP
#define XNTT_SVENTT_HPP

#include "kernel.hpp"
#include "layer.hpp"
#include "modmul.hpp"
#include "modulus.hpp"
#include "utility.hpp"
#include "vector.hpp"
#include "wrapper.hpp"

#endif
