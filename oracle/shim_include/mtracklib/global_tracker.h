// Forwarding header: what a maintainer puts in place of the reference's include/mtracklib/<this file>.
#pragma once
#include <rebvo_b200_shim.hpp>
