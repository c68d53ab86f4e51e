// Internal: friend gateway from the library into the kept classes' private state, plus the process-wide
// default handle the classes' GPU-backed predicates use.
#ifndef SC_ACCESS_H
#define SC_ACCESS_H

#include <string>

#include "sc_host.h"

class CascadeClassifier;

namespace sc_host {

struct Access {
    static bool flatten(CascadeClassifier& cc, int tmpl, FlatCascade* out, std::string* why);
};

// Lazily created handle on CUDA device 0; throws std::runtime_error when no GPU is usable (there is no CPU path).
sc_handle* default_handle();

}  // namespace sc_host

#endif
