/* TEST INFRASTRUCTURE stub for the MCU GPIO power markers used by evb/src/nnCntrlClass.c */
#ifndef NNSP_ORACLE_NS_ENERGY_STUB_H
#define NNSP_ORACLE_NS_ENERGY_STUB_H
#define NS_IDLE 0
#define NS_DATA_COLLECTION 1
#define NS_FEATURE_EXTRACTION 2
#define NS_INFERING 3
#define ns_set_power_monitor_state(x) ((void)0)
#endif
