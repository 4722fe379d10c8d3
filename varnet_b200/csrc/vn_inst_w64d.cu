// Class 164: hidden width <= 64, deep networks (5..8 hidden layers): 32-point tiles leave room for
// the weights of every layer in shared memory.
#define VN_CLS 164
#define VN_W 64
#define VN_TP_ADJ 32
#define VN_TP_FWD 32
#define VN_TP_RES 32
#define VN_TN 4
#include "vn_inst.cuh"
