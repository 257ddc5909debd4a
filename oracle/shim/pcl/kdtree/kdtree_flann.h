#include "../../pcl_shim.h"
