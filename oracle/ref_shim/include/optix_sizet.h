/* TEST INFRASTRUCTURE (oracle/_ref build only): stands in for the OptiX SDK 5.1 / CUDA header of this name, which is not
 * vendored under /root/reference.  Everything lives in oracle/ref_shim/dsref_runtime.h and dsref_vec.h. */
#include "../dsref_runtime.h"
