// placeholder, replaced below
#include "pio_host.h"
extern "C" int pio_attention_supported(int32_t dqk, int32_t dv) { return PIO_ERR_UNSUPPORTED; }
extern "C" int pio_attention_fwd(const pio_attention_args* a, void* stream) {
  return pio::fail(PIO_ERR_UNSUPPORTED, "pio_attention_fwd: not built yet");
}
