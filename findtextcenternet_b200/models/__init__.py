"""Host-side mirror of the reference's ``models`` package (models/detector.py, models/transformer.py):
same class names, constructor arguments, attribute paths and ``state_dict`` keys; ``forward`` runs on
the sm_100a engine behind include/ftc_b200.h."""
