/* TEST INFRASTRUCTURE stub for the MCU print service used by evb/src/nnCntrlClass.c and
 * PcmBufClass.c. Prints are routed to oracle/ref_glue.c, which uses the detection messages
 * to observe detections (the reference exposes them nowhere else). */
#ifndef NNSP_ORACLE_NS_HARNESS_STUB_H
#define NNSP_ORACLE_NS_HARNESS_STUB_H
void ref_glue_printf(const char *fmt, ...);
#define ns_lp_printf(...) ref_glue_printf(__VA_ARGS__)
#endif
