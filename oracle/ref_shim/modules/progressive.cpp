/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/progressive.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "progressive.cu"
#include "../dsref_device.h"
#include "CUDA/progressive.cu"
DSREF_BUFFER(frameResultBuffer)
DSREF_BUFFER(progressiveBuffer)
DSREF_BUFFER(varianceBuffer)
DSREF_PROGRAM(updateFrameResult)
DSREF_PROGRAM(clearScreen)
DSREF_PROGRAM(exception)
DSREF_PROGRAM(miss)
