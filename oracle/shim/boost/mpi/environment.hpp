// Stub of boost/mpi/environment.hpp for the oracle build (thread-backed ranks, see communicator.hpp).
#pragma once
namespace boost { namespace mpi { class environment { public: environment() {} template <class... A> environment(A&...) {} }; } }
