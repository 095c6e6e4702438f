// oracle/refshim — see opencv2/core.hpp (TEST INFRASTRUCTURE ONLY)
#include "core.hpp"
