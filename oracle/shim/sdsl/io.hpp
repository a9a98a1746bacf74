/* Stand-in header (oracle test infrastructure): see int_vector.hpp */
#include "int_vector.hpp"
