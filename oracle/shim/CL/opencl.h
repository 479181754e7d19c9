/* Empty stand-in: the reference's GPU test mains include <CL/opencl.h> but use nothing from it
 * (tests/test.exsum.gpu.cpp:14-18).  TEST INFRASTRUCTURE ONLY. */
