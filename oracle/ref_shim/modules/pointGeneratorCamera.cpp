/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/pointGeneratorCamera.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "pointGeneratorCamera.cu"
#include "../dsref_device.h"
#include "CUDA/pointGeneratorCamera.cu"
DSREF_BUFFER(directionBuffer)
DSREF_BUFFER(positionBuffer)
DSREF_PROGRAM(generatePoints)
DSREF_PROGRAM(clear)
DSREF_PROGRAM(exception)
DSREF_PROGRAM(miss)
