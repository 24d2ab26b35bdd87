// TEST INFRASTRUCTURE ONLY (oracle/): empty stand-in for a Boost.Serialization header.  The
// reference's serialize() member templates are never instantiated by the in-process MPI shim.
#pragma once
