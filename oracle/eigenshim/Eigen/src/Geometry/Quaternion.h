#include "../../Core"
