/*
 * varnet_b200.h — C ABI of the B200-native weak-form residual + gradient engine.
 *
 * This is the drop-in boundary for the reference's compute backend, the `TFNN`
 * object (`/root/reference/TFModel.py:54-437`) that `VarNet` holds as `tfData`
 * (`VarNet.py:201-203`) and drives through `sess.run(nodes, feed_dict)`
 * (`VarNetUtility.py:1044,1080,1086,1142`).  The reference has no FFI; the entry
 * points below are what a ctypes binding of that object needs (see INTEGRATION.md
 * and varnet_b200/_capi.py).  One handle = one GPU = one reference "tower"
 * (`TFModel.py:253-289`).
 *
 * Conventions: every function returns 0 on success or a negative VN_E_* code and
 * never throws; `vn_last_error` gives the message.  Host pointers are caller-owned
 * and only read/written during the call.  All work is enqueued on the handle's
 * stream (`vn_set_stream`); calls that return host data synchronise that stream.
 * Arithmetic is FP32 (the reference's placeholders are tf.float32,
 * TFModel.py:531,602-620); `*_f64` uploads round to FP32 exactly like the feed cast.
 */
#ifndef VARNET_B200_H
#define VARNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VN_MAX_LAYERS 8      /* hidden layers */
#define VN_MAX_INPDIM 8      /* dim + time + MOR parameters (VarNet.py:174-180) */

#define VN_ACT_SIGMOID 0     /* VarNet.py:162-164 default */
#define VN_ACT_TANH 1

#define VN_OPT_ADAM 0        /* tf.train.AdamOptimizer, TFModel.py:184 */
#define VN_OPT_RMSPROP 1     /* tf.train.RMSPropOptimizer, TFModel.py:186 */

#define VN_PROF_SLOTS 8
#define VN_MAX_TABLES 4096   /* resident point tables per engine (vn_select_table) */

#define VN_OK 0
#define VN_E_INVALID -1      /* bad argument (mirrors the reference's ValueError sites) */
#define VN_E_CUDA -2         /* CUDA runtime error */
#define VN_E_STATE -3        /* call order violated, e.g. loss before upload (VarNetUtility.py:1031) */
#define VN_E_UNSUPPORTED -4  /* network shape outside the compiled kernel families */

typedef struct vn_engine vn_engine;

/* Mirrors the TFNN constructor arguments that shape the graph (TFModel.py:85-86). */
typedef struct vn_config {
    int32_t dim;                       /* spatial dimension (1 or 2) */
    int32_t inpDim;                    /* MLP inputs: x, [t], MOR params (TFModel.py:517-518) */
    int32_t nLayers;                   /* number of hidden Dense layers */
    int32_t widths[VN_MAX_LAYERS];     /* layerWidth */
    int32_t act;                       /* VN_ACT_* (same for all hidden layers) */
    int32_t timeDependent;             /* TFModel.py:537,646,655 */
    int32_t isSource;                  /* lossOpt['isSource']   (TFModel.py:656) */
    int32_t integWflag;                /* lossOpt['integWflag'] (TFModel.py:660) */
    int32_t optimizer;                 /* VN_OPT_* */
    int32_t device;                    /* CUDA device ordinal */
} vn_config;

/* ---- lifetime ------------------------------------------------------------- */
int vn_create(const vn_config* cfg, vn_engine** out);
int vn_destroy(vn_engine* e);
const char* vn_last_error(void);
int vn_set_stream(vn_engine* e, void* cuda_stream);          /* cudaStream_t; NULL = default */
int vn_synchronize(vn_engine* e);

/* ---- parameters: flat Keras order, per layer kernel[in,out] row-major then bias;
 *      last is Dense(1) 'output' (TFModel.py:208-242).  Replaces tf.train.Saver /
 *      sess.run(tf.trainable_variables()) (VarNet.py:1362,2197-2239). */
int vn_param_count(const vn_engine* e, int64_t* n);
int vn_set_params(vn_engine* e, const float* theta, int64_t n);   /* also resets optimizer state */
int vn_get_params(vn_engine* e, float* theta, int64_t n);
int vn_get_optimizer_state(vn_engine* e, float* m, float* v, int64_t n, int64_t* step);
int vn_set_optimizer_state(vn_engine* e, const float* m, const float* v, int64_t n, int64_t step);

/* ---- per-tower feed (VarNetUtility.py:840-854).  Row-major host arrays exactly as
 *      the reference feeds them: Input[P,inpDim], gcoef[P,dim], source[P,1] or NULL,
 *      N[P,1] or NULL, dNt[P,1] or NULL; P = nb*integNum rows grouped test-function
 *      major.  integW[integNum] or NULL.  detJ: 1 value, or nb values when
 *      detJ_is_vector (the `detJvec` branch, TFModel.py:662-664).
 *      Data become device-resident; re-upload only when the caller changes them. */
int vn_upload_points_f32(vn_engine* e, const float* Input, const float* gcoef, const float* source,
                         const float* N, const float* dNt, int64_t nb, int32_t integNum,
                         const float* integW, const float* detJ, int32_t detJ_is_vector);
int vn_upload_points_f64(vn_engine* e, const double* Input, const double* gcoef, const double* source,
                         const double* N, const double* dNt, int64_t nb, int32_t integNum,
                         const double* integW, const double* detJ, int32_t detJ_is_vector);
/* biInput[nbi,inpDim], biLabel[nbi,1]; rows [0,bDof) are boundary rows, the rest
 * initial-condition rows (TFModel.py:643-650). */
int vn_upload_bic_f32(vn_engine* e, const float* biInput, const float* biLabel, int64_t nbi,
                      int64_t bDof, float biDimVal);
int vn_upload_bic_f64(vn_engine* e, const double* biInput, const double* biLabel, int64_t nbi,
                      int64_t bDof, double biDimVal);
/* ---- device-resident mini-batches.  The reference re-gathers `integInd[batchInd[n0:n1]]`
 *      on the host for every mini-batch / shuffle / MOR batch and re-feeds the copies
 *      (VarNetUtility.py:833-844,918-950,988-1013; VarNet.py:843-851).  Here several tables can stay
 *      resident (slots), a mini-batch is a list of test-function indices into the current table
 *      (on-device permutation; needs integNum % 4 == 0), and trailing MLP inputs that are constant over
 *      a batch (MOR parameters) are passed as scalars instead of re-tiled columns.
 *      vn_upload_table_*: like vn_upload_points_* but Input has only the first `nx` MLP input columns. */
/* vn_generate_table_f64: build the point table of a UNIFORM space-time mesh with CONSTANT coefficients on the device,
 * into the current table slot — what VarNet.trainingPoints (VarNet.py:576-586), PDEinpData/trainData
 * (gcoef = diff*dNx + vel*N, VarNet.py:837) and the periodic FE tables of FE.basisTot (FiniteElement.py:419-432)
 * produce on the host, without materialising or uploading nT rows.  For test function i = s*nTime + j (space index
 * slow, time index fast) and Gauss point q:
 *     Input[i,q,d] = coord[s][d] + hVec[d]*delta[d][q]  (d < dim),   Input[i,q,dim] = tcoord[j] + hVec[dim]*delta[dim][q]
 *     gcoef[i,q,k] = diff*dN[q][k] + vel[k]*N[q],   dNt[i,q] = dN[q][dim],   source*N = source*N[q]
 * in float64 with the host's operation order, rounded to float32 like the feed cast: the table is bit-identical to
 * uploading the host-built arrays.  Generates test functions [tf0, tf0+nb) (one tower's contiguous range).
 * coord [nSpace][dim], tcoord [nTime] (NULL when not time dependent), hVec [feDim], delta [feDim][integNum],
 * N [integNum], dN [integNum][feDim], vel [dim], integW [integNum] or NULL; detJ is the scalar Jacobian. */
int vn_generate_table_f64(vn_engine* e, const double* coord, int64_t nSpace, const double* tcoord, int64_t nTime,
                          const double* hVec, const double* delta, const double* N, const double* dN,
                          double diff, const double* vel, double source, int64_t tf0, int64_t nb,
                          int32_t integNum, const double* integW, double detJ);
int vn_select_table(vn_engine* e, int32_t slot);
int vn_table_loaded(const vn_engine* e, int32_t slot);                 /* 1 if the slot holds a table */
int vn_free_table(vn_engine* e, int32_t slot);
int vn_upload_table_f32(vn_engine* e, const float* Input, int32_t nx, const float* gcoef, const float* source,
                        const float* N, const float* dNt, int64_t nb, int32_t integNum,
                        const float* integW, const float* detJ, int32_t detJ_is_vector);
int vn_upload_table_f64(vn_engine* e, const double* Input, int32_t nx, const double* gcoef, const double* source,
                        const double* N, const double* dNt, int64_t nb, int32_t integNum,
                        const double* integW, const double* detJ, int32_t detJ_is_vector);
int vn_set_batch(vn_engine* e, const int32_t* tf_index, int64_t nb);   /* NULL: whole table, in order */
int vn_set_extra_inputs(vn_engine* e, const float* vals, int32_t n);

/* loss weights w[3] = [BC, IC, variational] (TFModel.py:666); device-resident so a
 * captured step graph sees updates. */
int vn_set_weights(vn_engine* e, const float w[3]);

/* ---- the hot path ----------------------------------------------------------
 * vn_loss:       forward only.  out[4] = {loss, BCloss, ICloss, varLoss}
 *                (TFModel.py:666,686-688); lossVec (nb floats, TFModel.py:668) may be NULL.
 * vn_loss_grad:  loss + d loss / d theta (NNModel.computeGrad, TFModel.py:695-714).
 *                Results stay on the device in the gradient buffer
 *                [grad(nparam) | loss, BCloss, ICloss, varLoss]; this is the buffer a
 *                multi-GPU caller all-reduces (SUM), mirroring TFNN.sum_grads
 *                (TFModel.py:342-377).  out may be NULL (no host sync).
 * vn_optimizer_step: TF-1.x Adam / RMSProp update from the gradient buffer
 *                (optimizer.apply_gradients, TFModel.py:313).
 * vn_train_step: loss_grad + optimizer_step in one call = one
 *                sess.run([optMinimize, loss]) (VarNetUtility.py:1044); single-GPU. */
int vn_loss(vn_engine* e, float out[4], float* lossVec);
int vn_loss_grad(vn_engine* e, float out[4]);
/* vn_loss_grad_fed_*: vn_upload_points_* + vn_loss_grad in one call, for callers that feed every array on
 * every step like the reference does (sess.run(..., feed_dict), VarNetUtility.py:1044; float64 -> float32
 * cast + copy per step): the table is uploaded chunk by chunk on a copy stream and the adjoint kernel is
 * launched per chunk as soon as it has been packed, so the copies overlap the step's kernels.  Same results
 * as the two separate calls; the caller's arrays are no longer read when the call returns.  Every kernel family takes
 * the chunks as they arrive (single-pass kernels per chunk; the two-pass class - integNum divides no tile - its forward
 * pass per chunk; the tensor-core class its own point chunks once the uploads cover them); tables of at most one
 * 4 Mi-row chunk are uploaded and then stepped. */
int vn_loss_grad_fed_f32(vn_engine* e, const float* Input, const float* gcoef, const float* source, const float* N,
                         const float* dNt, int64_t nb, int32_t integNum, const float* integW, const float* detJ,
                         int32_t detJvec, float out[4]);
int vn_loss_grad_fed_f64(vn_engine* e, const double* Input, const double* gcoef, const double* source, const double* N,
                         const double* dNt, int64_t nb, int32_t integNum, const double* integW, const double* detJ,
                         int32_t detJvec, float out[4]);
/* Page-lock a caller-owned host array for direct DMA (cudaHostRegister, portable) / undo it.  The reference re-feeds the SAME NumPy
 * arrays on every sess.run of an epoch loop (VarNetUtility.py:840-854 builds the dict once, :1044 runs it per step): once such an
 * array is registered, vn_loss_grad_fed_* / vn_upload_* copy straight out of it (float64 over the bus, cast by the pack kernel)
 * instead of casting it through pinned bounce buffers with host threads — which is what limits the fed step when several ranks
 * share the host's cores.  The caller keeps the memory alive until vn_host_unregister.  Process-wide, no engine needed.
 * vn_host_register returns 1 (not an error) when the range is already page-locked by its owner: do not unregister it. */
int vn_host_register(const void* ptr, size_t bytes);
int vn_host_unregister(const void* ptr);
int vn_grad_buffer(vn_engine* e, void** device_ptr, int64_t* n_floats);
int vn_get_grad(vn_engine* e, float* grad, int64_t n, float out[4]);
int vn_get_scalars(vn_engine* e, float out[4]);     /* {loss, BCloss, ICloss, varLoss} of the gradient buffer; synchronises */
/* lossVec[nb] = detJ_i R_i^2 (TFModel.py:668) as left by the last vn_loss / vn_loss_grad / vn_train_step of the
 * current batch (the fused adjoint kernels write it themselves); synchronises. */
int vn_get_lossvec(vn_engine* e, float* lossVec, int64_t nb);
/* Synchronises and reports a failure recorded by the asynchronous part of earlier calls (tensor-core classes: an
 * expired mbarrier wait; the step that hit it published NaN and skipped its optimizer update).  Clears the flag.
 * For callers that run steps without fetching the loss (vn_loss_grad(e, NULL), vn_train_step(e, lr, NULL)). */
int vn_check_error(vn_engine* e);
int vn_optimizer_step(vn_engine* e, float lr);
int vn_train_step(vn_engine* e, float lr, float* loss_out);
/* k steps on the current batch back to back (the captured step graph replayed k times, the k losses returned together):
 * k x sess.run([optMinimize, loss]) on an unchanged feed with one host round trip.  k <= 4096. */
int vn_train_steps(vn_engine* e, float lr, int32_t k, float* losses);
/* One optimizer step per mini-batch for k mini-batches of the current table with one host round trip: row i of
 * tf_index[k][nb] is the index list of step i, i.e. k x { vn_set_batch; vn_train_step } = one ManageTrainData.optimIter
 * over its mini-batches (VarNetUtility.py:1021-1047).  losses[k]; the engine is left on the last batch.  k <= 4096. */
int vn_train_batches(vn_engine* e, float lr, const int32_t* tf_index, int64_t nb, int32_t k, float* losses);
/* The same call in two halves: _begin copies the index lists into a pinned staging buffer and enqueues the copies, the k
 * steps and the read-back of their losses without waiting; _end waits for them and returns losses[k] (and reports an
 * expired tensor-core barrier wait).  Between the two the caller prepares the next feed dicts (the host work of the next
 * MOR parameter batch, VarNet.py:843-851) while the GPU runs.  Up to two calls may be in flight per engine (the next one
 * is enqueued, together with the vn_set_extra_inputs / vn_upload_bic_* that precede it, before the previous one is
 * collected: those small uploads are staged in pinned memory and do not wait either); _end collects the older one. */
int vn_train_batches_begin(vn_engine* e, float lr, const int32_t* tf_index, int64_t nb, int32_t k);
int vn_train_batches_end(vn_engine* e, float* losses, int32_t k);

/* ---- multi-GPU (one handle per GPU, one process per GPU): the towers' gradients and losses are summed like
 *      TFNN.sum_grads / optimSetup do on the controller (TFModel.py:315-319,342-377), with one NCCL all-reduce of the
 *      gradient buffer over NVLink.  NCCL is bound at run time (dlopen): `nccl_lib` may name the library, NULL uses the
 *      libnccl.so.2 already loaded into the process.  Protocol: tower 0 calls vn_comm_unique_id and ships the 128 bytes
 *      to the other processes (any host-side channel), then every tower calls vn_comm_init (collective).  From then on
 *      vn_train_step = kernels -> all-reduce -> optimizer update on the engine's stream, captured as ONE CUDA graph;
 *      vn_allreduce_grad is the same collective for callers that sequence vn_loss_grad / vn_optimizer_step themselves. */
int vn_comm_unique_id(const char* nccl_lib, void* id128);
int vn_comm_init(vn_engine* e, const char* nccl_lib, const void* id128, int32_t rank, int32_t world);
int vn_comm_world(const vn_engine* e);
int vn_allreduce_grad(vn_engine* e);

/* ---- evaluation (VarNetUtility.runSession, VarNetUtility.py:1098-1142)
 * vn_eval:     u = model(X), X[n,inpDim] host row-major, u[n].
 * vn_residual: strong-form residual  res = -u_t + kappa*Lap u - (vel - grad kappa).grad u + s
 *              (NNModel.Residual, TFModel.py:718-772).  diff[n], vel[n,dim],
 *              diff_dx[n,dim], source[n]; u and res are [n] (u may be NULL). */
int vn_eval_f32(vn_engine* e, const float* X, int64_t n, float* u);
int vn_eval_f64(vn_engine* e, const double* X, int64_t n, float* u);
int vn_residual_f64(vn_engine* e, const double* X, const double* diff, const double* vel,
                    const double* diff_dx, const double* source, int64_t n, float* u, float* res);

/* ---- measurement -----------------------------------------------------------
 * Per-kernel device time from CUDA events recorded on the engine's stream around each launch
 * (slots: 0 variational forward, 1 segmented reduce, 2 variational adjoint, 3 boundary/initial,
 * 4 finalize, 5 optimizer).  vn_profile_read synchronises, returns the totals since the last
 * read and resets them.  vn_fp32_peak_tflops runs an FFMA-only microbenchmark: the FP32
 * roofline denominator measured on the same GPU in the same run. */
int vn_profile_enable(vn_engine* e, int on);
int vn_profile_read(vn_engine* e, double ms[VN_PROF_SLOTS], int64_t counts[VN_PROF_SLOTS]);
int vn_fp32_peak_tflops(int device, int reps, double* tflops);

/* Development aid: cycle counters of the phases of the width-64 tensor-core tile kernel (CTA 0, one thread) for the
 * last launch; filled only when the environment variable VARNET_B200_TC64_TIMING is set.  See scripts/tc64_phases.py. */
int vn_debug_tc64_timing(int64_t out[16]);

/* ---- introspection for tests / bench --------------------------------------- */
int vn_kernel_info(const vn_engine* e, char* buf, size_t buflen);   /* kernel family, tile, smem */
int64_t vn_launch_count(const vn_engine* e);                         /* kernels launched so far */

#ifdef __cplusplus
}
#endif
#endif /* VARNET_B200_H */
