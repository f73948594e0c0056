// Forwarding header of the oracle-only abseil stand-in (see absl_shim_all.h).
#include "open_spiel/abseil-cpp/absl/absl_shim_all.h"
