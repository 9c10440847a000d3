"""ORACLE — test infrastructure only (see each module's header).  The product package
`unsupervised_domain_adaptation_object_detection_implementation_b200` never imports this."""
