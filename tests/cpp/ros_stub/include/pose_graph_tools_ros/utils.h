/* pose_graph_tools_ros/utils.h of the ROS stand-in: dpgo_ros includes it (src/PGOAgentROS.cpp:15) and uses nothing from it. */
