#include "geometry_msgs/Point.h"
