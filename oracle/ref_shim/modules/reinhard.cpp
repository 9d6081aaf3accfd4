/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/reinhard.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "reinhard.cu"
#include "../dsref_device.h"
#include "CUDA/reinhard.cu"
DSREF_BUFFER(progressiveBuffer)
DSREF_BUFFER(varianceBuffer)
DSREF_BUFFER(sumLuminanceColumns)
DSREF_BUFFER(averageLuminance)
DSREF_BUFFER(screenBuffer)
DSREF_PROGRAM(firstPass)
DSREF_PROGRAM(secondPass)
DSREF_PROGRAM(applyReinhard)
