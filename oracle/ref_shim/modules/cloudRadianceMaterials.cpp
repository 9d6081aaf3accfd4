/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/cloudRadianceMaterials.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "cloudRadianceMaterials.cu"
#include "../dsref_device.h"
#include "CUDA/cloudRadianceMaterials.cu"
DSREF_SAMPLER(inScatter)
DSREF_SAMPLER(mie)
DSREF_SAMPLER(choppedMie)
DSREF_SAMPLER(choppedMieIntegral)
DSREF_PROGRAM(totalRadiance)
DSREF_PROGRAM(multipleScatterSunRadiance)
DSREF_PROGRAM(singleScatterSunRadiance)
