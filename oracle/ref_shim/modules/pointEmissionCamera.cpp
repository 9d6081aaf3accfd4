/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/pointEmissionCamera.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "pointEmissionCamera.cu"
#include "../dsref_device.h"
#include "CUDA/pointEmissionCamera.cu"
DSREF_BUFFER(tasks)
DSREF_PROGRAM(estimateEmission)
DSREF_PROGRAM(clear)
