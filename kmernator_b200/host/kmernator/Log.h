// Log.h -- the small part of the reference's logging surface the hot-path drivers use
// (src/Log.h:431-456,484): LOG_VERBOSE / LOG_DEBUG / LOG_WARN / LOG_ERROR / LOG_THROW and LoggedException.
// Plain stderr; no MPI gathering (out of scope, SURVEY.md section 2.1 "Log").
#ifndef KMERNATOR_HOST_LOG_H
#define KMERNATOR_HOST_LOG_H

#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>

class LoggedException : public std::runtime_error {
public:
    explicit LoggedException(const std::string &msg) : std::runtime_error(msg) {}
};

class Log {
public:
    static int &verboseLevel() { static int v = 1; return v; }
    static int &debugLevel() { static int d = 0; return d; }
    static bool isVerbose(int level) { return verboseLevel() >= level; }
    static bool isDebug(int level) { return debugLevel() >= level; }
};

#define LOG_VERBOSE(level, msg) do { if (Log::isVerbose(level)) { std::cerr << msg << std::endl; } } while (0)
#define LOG_VERBOSE_OPTIONAL(level, cond, msg) do { if ((cond) && Log::isVerbose(level)) { std::cerr << msg << std::endl; } } while (0)
#define LOG_DEBUG(level, msg) do { if (Log::isDebug(level)) { std::cerr << "DEBUG" << level << ": " << msg << std::endl; } } while (0)
#define LOG_WARN(level, msg) do { std::cerr << "WARNING: " << msg << std::endl; } while (0)
#define LOG_ERROR(level, msg) do { std::cerr << "ERROR: " << msg << std::endl; } while (0)
#define LOG_THROW(msg) do { std::ostringstream ss_; ss_ << msg; throw LoggedException(ss_.str()); } while (0)

#endif
