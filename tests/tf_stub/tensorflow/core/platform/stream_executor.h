// TEST INFRASTRUCTURE, NOT TENSORFLOW: see ../framework/op_kernel.h.
