#pragma once
// stand-in for the traits the functor headers specialise (exaNBody compute_cell_particles.h)
namespace exanb { template <class F> struct ComputeCellParticlesTraits { static inline constexpr bool RequiresBlockSynchronousCall = false; static inline constexpr bool CudaCompatible = false; }; }
