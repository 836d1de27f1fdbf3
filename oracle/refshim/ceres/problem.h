// forwards to the oracle stand-in (oracle/refshim/lorb_ceresshim.hpp); test infrastructure only
#include "../lorb_ceresshim.hpp"
