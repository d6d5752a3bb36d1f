/* TEST STAND-IN, see pg_nodes_stub.h */
#include "../pg_nodes_stub.h"
