// Width class 32 (the reference's operator configs: [20], [10,20], [10,20,30]).
#define VN_W 32
#define VN_TP_ADJ 64
#define VN_TP_FWD 64
#define VN_TN 4
#include "vn_inst.cuh"
