// Class 16: hidden width <= 16.  256-point tiles (256 threads: 64 point groups x 4 neuron groups), so the narrow
// layers are not padded to the 32-wide class (4x fewer hidden-layer FMAs than class 32 at width 16).
#define VN_CLS 16
#define VN_W 16
#define VN_TP_ADJ 256
#define VN_TP_FWD 256
#define VN_TP_RES 128
#define VN_TN 4
#include "vn_inst.cuh"
