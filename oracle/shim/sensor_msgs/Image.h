/* stub header: everything lives in oracle/shim/ros_shim.h (test infrastructure, see README there) */
#include "../ros_shim.h"
