// placeholder translation unit: replaced by the tcgen05 flash-attention kernel
#include "odise_b200.h"
extern "C" int odise_attention_tc(const void*, const void*, long long, const void*, const void*, long long,
                                  const void*, const void*, long long, long long, float*, void*, void*, long long,
                                  int, int, int, int, int, float, int, void*) {
  return ODISE_ERR_UNSUPPORTED;
}
