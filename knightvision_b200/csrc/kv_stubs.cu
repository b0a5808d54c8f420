// temporary: net / mcts teardown hooks until those translation units exist
#include "kv_internal.h"
void kv_mcts_destroy(kv_ctx*) {}
