"""onsas.jl_b200 -- B200-native Newton-Raphson hot path behind ONSAS.jl's API."""
