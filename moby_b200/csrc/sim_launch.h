// Kernel entry points of the stepped path, one per translation unit; the host side (sim_kernels.cu) launches them
// through cudaLaunchKernel.  Argument lists are documented next to each kernel.
#pragma once
const void* b2m_k_step_warp();            // (SimParams P, double dt, int n_steps, size_t env_d)
const void* b2m_k_finish();               // (SimParams P, double dt, int round)
const void* b2m_k_advance();              // (SimParams P, double dt, int round, int wpb)
const void* b2m_k_advance_thread(int cls); // (SimParams P, double dt, int round): thread per env; cls 0/1/2 = local working set of 256/1024/4096 doubles
// thread-per-env impact kernels: variant v keeps a working set of up to B2M_THREAD_ND<v> doubles / B2M_THREAD_NI<v> ints in local memory
#define B2M_THREAD_ND0 2048
#define B2M_THREAD_NI0 288
#define B2M_THREAD_ND1 5632
#define B2M_THREAD_NI1 384
const void* b2m_k_impact_thread(int variant);   // (SimParams P, double dt, int round, int slot)
const void* b2m_k_impact_warp();          // (SimParams P, double dt, int round, int slot, int wpb, LadderPool L, int feed_slot, int* feed_done, int feed_expect)
const void* b2m_k_impact_subwarp8();       // (SimParams P, double dt, int round, int slot, int wpb): eight lanes per env, 4 * wpb envs per block
const void* b2m_k_signal();               // (int* counter): one thread, counter += 1 (stream-ordered completion signal of a class launch)
const void* b2m_k_impact_block64();
const void* b2m_k_impact_block128();
const void* b2m_k_impact_block256();
const void* b2m_k_finish_block256();      // (SimParams P, double dt, int round): block-per-env finish for large-LCP scenes
const void* b2m_k_rc_fwd_dyn();           // (SimParams P, int algo, const double* jq, const double* jqd, const double* tau, double* qdd)
const void* b2m_k_rc_inertia();           // (SimParams P, const double* jq, double* H)
const void* b2m_k_rc_refresh();           // (SimParams P)
// constraint stabilization: thread per env with a local working set of up to B2M_STAB_ND<v> doubles / B2M_STAB_NI<v> ints, else warp per env
#define B2M_STAB_ND0 1024
#define B2M_STAB_NI0 192
#define B2M_STAB_ND1 3072
#define B2M_STAB_NI1 512
const void* b2m_k_stabilize_thread(int variant);   // (SimParams P, int mode, int* queue, int* count) with P.nmax = P.cmax; mode 0: select envs into queue, 1: stabilize the queued envs (queue == nullptr: all)
const void* b2m_k_stabilize_warp();               // (SimParams P, size_t nd_env, size_t nd_all, int* queue, int* count): queue == nullptr: every env
