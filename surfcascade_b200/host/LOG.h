// Compile-time levelled logging to std::cout, same macro names as the reference's LOG.h:10-24.
#ifndef LOG_H
#define LOG_H

#include <iostream>

enum LEVELS { ERROR_LEVEL, WARNING_LEVEL, INFO_LEVEL, DEBUG_LEVEL };

#ifndef SETLEVEL
#define SETLEVEL INFO_LEVEL
#endif

#define SC_LOG_AT(LVL, MSG, TAIL) do { if (SETLEVEL >= (LVL)) std::cout << MSG TAIL; } while (0)
#define LOG_DEBUG(MSG)      SC_LOG_AT(DEBUG_LEVEL, MSG, << std::endl)
#define LOG_INFO(MSG)       SC_LOG_AT(INFO_LEVEL, MSG, << std::endl)
#define LOG_WARNING(MSG)    SC_LOG_AT(WARNING_LEVEL, MSG, << std::endl)
#define LOG_ERROR(MSG)      SC_LOG_AT(ERROR_LEVEL, MSG, << std::endl)
#define LOG_DEBUG_NN(MSG)   SC_LOG_AT(DEBUG_LEVEL, MSG, )
#define LOG_INFO_NN(MSG)    SC_LOG_AT(INFO_LEVEL, MSG, )
#define LOG_WARNING_NN(MSG) SC_LOG_AT(WARNING_LEVEL, MSG, )
#define LOG_ERROR_NN(MSG)   SC_LOG_AT(ERROR_LEVEL, MSG, )

#endif
