"""The reference's process functions (transforms/*.py), same call contract
``f(*input_paths, output_dirs, **options) -> Path | List[Path] | None``, GPU arithmetic."""
