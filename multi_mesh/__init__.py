"""Drop-in alias package: `import multi_mesh...` resolves to the B200-native implementation
(multimesh_b200).  Module paths match the reference's layout."""
