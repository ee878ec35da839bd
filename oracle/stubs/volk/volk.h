/* Oracle stub for <volk/volk.h>: the reference only asks for the alignment. */
#ifndef ORACLE_STUB_VOLK_H
#define ORACLE_STUB_VOLK_H
#include <stddef.h>
static inline size_t volk_get_alignment(void) { return 32; }
#endif
