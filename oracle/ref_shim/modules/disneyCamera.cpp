/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/disneyCamera.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "disneyCamera.cu"
#include "../dsref_device.h"
#include "CUDA/disneyCamera.cu"
DSREF_BUFFER(networkInputBuffer)
DSREF_BUFFER(directRadianceBuffer)
DSREF_BUFFER(predictedRadianceBuffer)
DSREF_BUFFER(frameResultBuffer)
DSREF_PROGRAM(pinholeCamera)
DSREF_PROGRAM(copyToFrameResult)
DSREF_PROGRAM(clearRect)
