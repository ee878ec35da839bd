/* Oracle stub for <gnuradio/io_signature.h> (test infrastructure). */
#ifndef ORACLE_STUB_GR_IO_SIGNATURE_H
#define ORACLE_STUB_GR_IO_SIGNATURE_H
#include <gnuradio/block.h>
#endif
