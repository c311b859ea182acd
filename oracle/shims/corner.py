"""Import stub for mc_plot's `import corner`."""
