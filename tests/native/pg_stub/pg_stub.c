/* TEST STAND-IN, see postgres.h in this directory */
#include "postgres.h"
char pg_stub_error_message[256];
int  pg_stub_error_code;
