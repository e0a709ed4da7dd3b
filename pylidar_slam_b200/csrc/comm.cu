// placeholder until the NCCL glue lands (K9)
#include "internal.cuh"
namespace pls {
int comm_rank(pls_context*) { return 0; }
int comm_size(pls_context*) { return 1; }
void comm_allreduce_sums(pls_context*, double*) {}
void comm_free(pls_context*) {}
}
extern "C" {
int pls_comm_init(pls_context*, int, int, const void*, const char*) { return PLS_E_COMM; }
int pls_comm_unique_id(const char*, void*) { return PLS_E_COMM; }
int pls_comm_destroy(pls_context*) { return PLS_OK; }
}
