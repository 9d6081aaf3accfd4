/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/inScatter.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "inScatter.cu"
#include "../dsref_device.h"
#include "CUDA/inScatter.cu"
DSREF_BUFFER(inScatterBuffer)
DSREF_SAMPLER(density)
DSREF_PROGRAM(inScatter)
