/* ort_internal.h -- declarations shared by the translation units of libort.so */
#ifndef ORT_INTERNAL_H
#define ORT_INTERNAL_H

#include "../../include/ort.h"
#include "ort_dev_types.h"

#if defined(__GNUC__)
__attribute__((format(printf, 1, 2)))
#endif
void ort_set_error(const char* fmt, ...);
/* CUDA device index of the library's first device, or -1 when it is not initialised */
int ort_internal_primary_device(void);


#endif
