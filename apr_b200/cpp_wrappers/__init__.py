"""Drop-in replacements for Predator_APR/cpp_wrappers (same module paths, function names and keyword signatures)."""
