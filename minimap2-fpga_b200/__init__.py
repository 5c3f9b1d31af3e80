"""B200-native chaining backend for minimap2 (mm_chain_dp hot path). See DESIGN.md."""
