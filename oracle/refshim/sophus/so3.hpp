#include "se3.hpp"
