// Programmatic dependent launch (griddepcontrol) helpers.  Both instructions are no-ops in a grid that was not launched
// through a programmatic dependency, so the kernels that carry them run unchanged on plain stream launches.
//   pdl_launch_dependents: the grids that depend on this one programmatically may be scheduled as soon as every CTA of this
//                          grid has executed it (or exited); their CTAs become resident and block in pdl_wait.
//   pdl_wait:              returns once every prerequisite grid has completed and its memory operations are visible.
//                          Data written by a prerequisite grid must be read with coherent loads after it (no ld.global.nc).
// Used by the k-steps-in-one-graph path of the launch-bound operator configurations (vn_capi.cu: pdl_edges).
#pragma once
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
