#include "../boost_shim.h"
