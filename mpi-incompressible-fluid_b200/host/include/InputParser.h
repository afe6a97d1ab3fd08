// InputParser.h -- drop-in for include/InputParser.h:11-14: the `key : value` input file of `mif`.
#ifndef INPUT_PARSER_H
#define INPUT_PARSER_H

#include <cstddef>
#include <string>

#include "Real.h"

namespace mif {

// Required keys: Nx, Ny, Nz, dt, Nt, Py, Pz, test_case_2.  Throws std::runtime_error on an unreadable file, an
// unknown or repeated key, or a missing key (src/InputParser.cpp:15-76).
void parse_input_file(const std::string &filename, size_t &Nx_global, size_t &Ny_global, size_t &Nz_global, Real &dt,
                      unsigned int &num_time_steps, int &Py, int &Pz, bool &test_case_2);

}  // namespace mif

#endif  // INPUT_PARSER_H
