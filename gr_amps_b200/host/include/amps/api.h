// api.h -- export macro, as include/amps/api.h:6-10 of the reference
#pragma once
#if defined(__GNUC__)
#define AMPS_API __attribute__((visibility("default")))
#else
#define AMPS_API
#endif
