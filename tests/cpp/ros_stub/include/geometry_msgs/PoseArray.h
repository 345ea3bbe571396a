#include "geometry_msgs/Pose.h"
