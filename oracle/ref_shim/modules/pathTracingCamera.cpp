/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/pathTracingCamera.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "pathTracingCamera.cu"
#include "../dsref_device.h"
#include "CUDA/pathTracingCamera.cu"
DSREF_BUFFER(frameResultBuffer)
DSREF_PROGRAM(pinholeCamera)
DSREF_PROGRAM(miss)
