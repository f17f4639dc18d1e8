// placeholder; replaced by the tcgen05 backward
#include "../../include/b200_ffa.h"
extern "C" {
size_t b200_ffa_bwd_workspace_bytes(const b200_ffa_bwd_params*) { return 0; }
int b200_ffa_bwd(const b200_ffa_bwd_params*, b200_stream_t) { return B200_ERR_UNSUPPORTED; }
}
