// kv_tables_dev.cuh — the attack tables as a per-translation-unit device constant (no -rdc needed);
// kernels stage them into shared memory once per CTA.
#pragma once
#include "kv_tables.cuh"
namespace kv {
static __device__ const Tables g_tables = make_tables();
}
