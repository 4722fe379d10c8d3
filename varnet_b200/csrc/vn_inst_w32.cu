// Class 32: hidden width <= 32 (the reference's operator configs: [20], [10,20], [10,20,30]).
// 128-point tiles, 256 threads.
#define VN_CLS 32
#define VN_W 32
#define VN_TP_ADJ 128
#define VN_TP_FWD 128
#define VN_TP_RES 64
#define VN_TN 4
#include "vn_inst.cuh"
