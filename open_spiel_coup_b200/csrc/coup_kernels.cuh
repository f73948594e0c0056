// coup_kernels.cuh -- every sm_100a kernel of the batched Coup environment, by topic:
//   coup_device.cuh       the rules on the packed per-env state (branch-free step, legal masks, Philox, observer head)
//   coup_step.cuh         env slab, per-env decision step, reset / step / single-move / fork kernels, statistics, episode ring
//   coup_policy.cuh       uniform and masked-softmax action sampling, dense legal mask
//   coup_encode.cuh       info-state and observation encoders (staged + bulk stores, plain stores), row hash
//   coup_rollout.cuh      fused rollout step kernels (k_rollout_ws and variants, env-only multi-step)
//   coup_incremental.cuh  incremental info-state contract
//   coup_record.cuh       reservoir + replay recording inside the step
//   coup_cfr.cuh          levels of sampled CFR traversals
#pragma once
#include "coup_step.cuh"
#include "coup_policy.cuh"
#include "coup_encode.cuh"
#include "coup_rollout.cuh"
#include "coup_incremental.cuh"
#include "coup_record.cuh"
#include "coup_cfr.cuh"
