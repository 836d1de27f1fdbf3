// forwards to the oracle stand-in (oracle/refshim/lorb_cvshim.hpp); test infrastructure only
#include "../../lorb_cvshim.hpp"
