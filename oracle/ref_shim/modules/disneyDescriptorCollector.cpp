/* TEST INFRASTRUCTURE (oracle/_ref build only): compiles the reference's CUDA/disneyDescriptorCollector.cu, unmodified and where it lies under
 * /root/reference (-I $(REF_SRC)), as one emulated PTX module; see ../dsref_device.h. */
#define DSREF_MODULE_NAME "disneyDescriptorCollector.cu"
#include "../dsref_device.h"
#include "CUDA/disneyDescriptorCollector.cu"
DSREF_BUFFER(descriptors)
DSREF_BUFFER(directionBuffer)
DSREF_BUFFER(positionBuffer)
DSREF_PROGRAM(collect)
DSREF_PROGRAM(clear)
