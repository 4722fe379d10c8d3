// Class 64: hidden width <= 64.  64-point tiles, 256 threads; three 52 KB operand buffers + the
// weights of up to 4 hidden layers fit the 227 KB of shared memory of one SM.
#define VN_CLS 64
#define VN_W 64
#define VN_TP_ADJ 64
#define VN_TP_FWD 64
#define VN_TP_RES 32
#define VN_TN 4
#include "vn_inst.cuh"
