// Width class 64: adjoint tiles of 32 points keep all layers of a 4x64 dim-2 network in
// shared memory (~218 KB); forward-only tiles use 64 points with ping-pong buffers.
#define VN_W 64
#define VN_TP_ADJ 32
#define VN_TP_FWD 64
#define VN_TN 4
#include "vn_inst.cuh"
