/* Oracle stub (test infrastructure, not product code): stands in for
 * <gnuradio/attributes.h> so the reference block sources compile unchanged. */
#ifndef ORACLE_STUB_GR_ATTRIBUTES_H
#define ORACLE_STUB_GR_ATTRIBUTES_H
#define __GR_ATTR_EXPORT __attribute__((visibility("default")))
#define __GR_ATTR_IMPORT __attribute__((visibility("default")))
#endif
