// Shadows the reference's open_spiel/games/coup.h on the include path (plugin/ comes first), so that the
// reference's own games/coup_test.cc compiles UNMODIFIED against the GPU-backed plugin classes.
#include "coup_b200_plugin.h"
