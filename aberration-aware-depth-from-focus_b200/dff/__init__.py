"""Drop-in mirror of the parts of the reference's ``dff`` package that sit next to the synthesis path."""
