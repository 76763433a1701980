"""Decoy sibling: stands where the reference's own multi_input_vocoder/models_multi_input.py stands next to
inference_server.py.  If `from models_multi_input import MelCodeGenerator` ever resolves to THIS file, the drop-in
launcher failed to take precedence over the script's directory."""
raise ImportError("decoy models_multi_input.py imported: the drop-in launcher did not win over sys.path[0]")
