"""reconvat_b200 -- B200-native Mel front-end + VAT perturbation loop."""
