"""quantum_systems_b200 -- B200-native two-body integral pipeline behind the HyQD/quantum-systems API."""

__version__ = "0.1.0"
