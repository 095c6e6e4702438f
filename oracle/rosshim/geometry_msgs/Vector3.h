#include "../rosshim.hpp"
