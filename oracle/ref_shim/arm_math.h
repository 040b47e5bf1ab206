/* empty: evb/src/nnCntrlClass.c:2 includes <arm_math.h> "for fft only" and uses nothing from it */
